# A/B timing of the TMA-staged depthwise kernel (tile budget / stages)
timeout 200 python -m pytest tests -m gpu -q -x -k "stagewise or golden or strict or stock" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || exit 1
for cfg in ${DW_CFGS:-"48 2" "32 2" "32 3"}; do set -- $cfg; echo "tile_kb=$1 stages=$2"; KPD_DW_TILE_KB=$1 KPD_DW_STAGES=$2 timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernels_ms']['body'])"; done
bash tools/ncu_list.sh | grep -E "dw_tma|launches|stem|se_fc"
