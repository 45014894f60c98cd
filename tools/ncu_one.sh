# full ncu metrics of one kernel launch: KREGEX, KSKIP
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${KSKIP:-3} -c 1 -o gpurun_out/prof_one -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_one.log 2>&1
ncu -i gpurun_out/prof_one.ncu-rep --page raw --csv > gpurun_out/ncu_one_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_one.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_one_sass.csv 2>/dev/null
rm -f gpurun_out/prof_one.ncu-rep
