# launch list of one bench step (cold-cache, serialised per-launch times): all launches are captured, the last full step is printed
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_all.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
body=[r for r in rows[hi+1:] if len(r)>vi]
starts=[i for i,r in enumerate(body) if r[ki].startswith("stem_kernel")]
a,b=starts[-2],starts[-1]
step=body[a:b]
with open("gpurun_out/launches_step.csv","w",newline="") as f:
    w=csv.writer(f); w.writerow(h); w.writerows(step)
tot=0
for r in step:
    t=float(r[vi].replace(",",""))/1000; tot+=t
    print(r[ki][:34].ljust(34), r[gi].rjust(16), round(t,1))
print("launches", len(step), "sum_us", round(tot,1))
PY
rm -f gpurun_out/launches_all.csv
