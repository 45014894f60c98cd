python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 70 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_v3.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
ki=h.index("Kernel Name"); vi=h.index("Metric Value"); gi=h.index("Grid Size")
for r in rows[hi+1:]:
    if len(r)>vi: print(r[ki][:28].ljust(28), r[gi].rjust(16), round(float(r[vi].replace(",",""))/1000,1))
PY
