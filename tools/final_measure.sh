# End-of-round measurement set (single GPU): tests, smoke, the default bench line (C2 bf16 + fp32 / C5 / C4 sub-records + the CPU
# baseline), the reference arm and the P=3 variant.  JSON lines land in gpurun_out/.
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 1200 python bench.py 2>gpurun_out/bench_bf16.err | tail -1 > gpurun_out/bench_bf16.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
timeout 300 python bench.py --steps 20 --warmup 3 --persons 3 --sub '' --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_p3_bf16.json
for f in bench_bf16 bench_reference bench_p3_bf16; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e)
PY
done
