# End-of-round measurement set (single GPU): tests, smoke, bench in both precisions with the CPU baseline, the reference arm,
# the multi-person 512x512 configuration (C5) and the decode / NMS micro-benchmarks.  JSON lines land in gpurun_out/.
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 900 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_bf16.err | tail -1 > gpurun_out/bench_bf16.json
timeout 300 python bench.py --steps 50 --warmup 5 --precision fp32 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_fp32.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
timeout 300 python bench.py --steps 10 --warmup 3 --batch 128 --size 512 --persons 3 --no-cpu-baseline 2>gpurun_out/bench_c5.err | tail -1 > gpurun_out/bench_c5_bf16.json
timeout 300 python bench.py --steps 20 --warmup 3 --persons 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_p3_bf16.json
timeout 600 python tools/bench_micro.py > gpurun_out/micro.json 2>gpurun_out/micro.err
for f in bench_bf16 bench_fp32 bench_reference bench_c5_bf16 bench_p3_bf16; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json"))
    print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e)
PY
done
tail -5 gpurun_out/micro.json
