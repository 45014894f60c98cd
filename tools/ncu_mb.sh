for sk in 0 2 1; do
KPD_MB_SKIP=$sk ncu --metrics gpu__time_duration.sum --clock-control none -k regex:mbconv -s 33 -c 11 --csv --log-file gpurun_out/mb_skip$sk.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/mb_skip$sk.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); h=rows[hi]
print("skip=$sk", [round(float(r[h.index("Metric Value")].replace(",",""))/1000) for r in rows[hi+1:] if len(r)>5])
PY
done
