# ncu --set full of ONE launch: KREGEX (kernel name regex), KSKIP (matching launches to skip; default 3 warm-up steps x PER launches per step),
# TAG for the output names.  Leaves the raw metric page and the SASS source page as CSV (the .ncu-rep stays on the box).
TAG=${TAG:-one}
python tools/step_once.py $ARGS > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${KSKIP:-3} -c 1 -o gpurun_out/${TAG}_prof -f python tools/step_once.py $ARGS > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
rm -f gpurun_out/${TAG}_prof.ncu-rep
