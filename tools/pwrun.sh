# A/B: CUDA-core FFMA pointwise GEMM (default) vs the 3xTF32 mma.sync predecessor (KPD_PW_MMA=1)
timeout 200 python -m pytest tests -m gpu -q -x -k "stagewise or golden or strict or stock" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
for v in "" "KPD_PW_MMA=1"; do echo "variant: $v"; env $v timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels_ms']['body'], d['kernels_ms']['fpn_laterals'])"; done
bash tools/ncu_list.sh | grep -E "pw_|launches"
