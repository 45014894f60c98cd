"""bench.py contract checks that run without a GPU: the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "images/sec" and line["value"] > 0
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dll", "models"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_native_arm_does_not_touch_the_oracle_outside_the_cpu_baseline():
    src = open(os.path.join(ROOT, "bench.py")).read()
    head, tail = src.split("def cpu_reference", 1)
    body_of_cpu_ref, rest = tail.split("\ndef ", 1)
    assert "oracle" not in head.split('"""', 2)[-1].replace("oracle/", "")      # nothing before cpu_reference imports it
    assert "import oracle" not in rest and "from oracle" not in rest
