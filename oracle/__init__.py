"""CPU oracle for the keypoint-detection inference hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is an independent CPU restatement (torch-CPU fp32 functional ops + numpy) of
the reference's `MultiPersonKeypointModel` inference path, written from the
reference's behaviour (every function cites the reference file:line it follows,
paths relative to the upstream repo root; `tv:` = torchvision 0.26.0).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it -- and only as the checker / the CPU
baseline, never as the thing shipped.  Nothing under `keypoint-detection_b200/`
imports this package.

Pinning: the reference ships NO tests and NO golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, run in the build container by `tests/golden/make_golden.py` (which
imports `/root/reference/dll` unmodified apart from the `weights=None` patch)
and committed as `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks the
oracle against those vectors.
"""
from .arch import MBV3_SMALL_BLOCKS, state_dict_spec          # noqa: F401
from .weights import synth_state_dict, synth_images, synth_boxes  # noqa: F401
