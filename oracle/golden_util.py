"""Shared helpers for golden fixtures (test infrastructure)."""
import numpy as np


def sample_big(t, max_elems=40000):
    """Deterministic strided sub-sample of a big NCHW tensor (keeps fixtures small).

    Spatial dims are strided by primes (7, 5) so rows/cols of every residue are hit across channels.
    """
    a = np.asarray(t, dtype=np.float32)
    if a.size <= max_elems or a.ndim != 4:
        return a
    s = a[:, :, ::7, ::5]
    if s.size > max_elems:
        step = int(np.ceil(s.size / max_elems))
        s = s[:, ::step]
    return np.ascontiguousarray(s)


def checksum(t):
    """Order-independent float64 moments, robust to 1e-6 relative noise at tolerance 1e-4."""
    a = np.asarray(t, dtype=np.float64)
    return np.asarray([a.sum(), np.abs(a).sum(), (a * a).sum()], dtype=np.float64)


# canonical golden cases: name -> dict(batch, size, persons, zero_slots)
GOLDEN_CASES = {
    "s224_b2_p3": dict(batch=2, size=224, persons=3, zero_slots=((0, 1),)),
    "s224_b2_p2_empty_image": dict(batch=2, size=224, persons=2, zero_slots=((0, 0), (0, 1))),
    "s512_b1_p2": dict(batch=1, size=512, persons=2, zero_slots=()),
}
