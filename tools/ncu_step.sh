# Per-round profile of ONE forward step (C2 bf16 unless ARGS says otherwise): (1) plain run, (2) launch list with
# gpu__time_duration, (3) ncu --set full metrics of every launch of the step.  Only CSV extracts come back.
#   TAG=r02 ARGS="--precision bf16" bash tools/ncu_step.sh
TAG=${TAG:-step}
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum"
python tools/step_once.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 || { cat gpurun_out/${TAG}_plain.log; exit 1; }
NL=$(grep -oE 'launches_per_step [0-9]+' gpurun_out/${TAG}_plain.log | grep -oE '[0-9]+')
echo "launches per step: $NL"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * NL)) -c $NL --csv --log-file gpurun_out/${TAG}_launches.csv python tools/step_once.py $ARGS > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${TAG}_launches.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
tot=0; out=[]
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    t=float(r[vi].replace(",",""))/1000; tot+=t
    out.append(f"{r[ki][:60].ljust(60)} {r[gi].rjust(16)} {t:8.1f}")
out.append(f"launches {len(out)} sum_us {tot:.1f}")
open("gpurun_out/${TAG}_launches.txt","w").write("\n".join(out)+"\n")
print(out[-1])
PY
if [ "${FULL:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -s $((3 * NL)) -c $NL -o gpurun_out/${TAG}_prof -f python tools/step_once.py $ARGS > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv --metrics $M > gpurun_out/${TAG}_ncu_metrics.csv 2>/dev/null
[ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/${TAG}_prof.ncu-rep
fi
