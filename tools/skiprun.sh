run() { echo -n "$* -> "; env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['kernels_ms']['body'],3))"; }
for sk in 0 1 2 4 8 15; do run KPD_MB_SKIP=$sk; done
run KPD_MBCONV_FMA=1 KPD_MB_SKIP=0; run KPD_MBCONV_FMA=1 KPD_MB_SKIP=1
