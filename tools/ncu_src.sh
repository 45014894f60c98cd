# source-level stall attribution of one kernel launch: KREGEX (kernel name regex), KSKIP (matching launches to skip)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-pw_ffma}" -s ${KSKIP:-7} -c 1 -o gpurun_out/prof_src -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1
ncu -i gpurun_out/prof_src.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_src_sass.csv 2>/dev/null
ncu -i gpurun_out/prof_src.ncu-rep --page raw --csv > gpurun_out/ncu_src_raw.csv 2>/dev/null
rm -f gpurun_out/prof_src.ncu-rep
ls -la gpurun_out/ncu_src*
