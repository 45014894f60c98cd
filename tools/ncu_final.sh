# Final per-round profile: (1) plain run, (2) launch list of one step, (3) ncu --set full of every launch of one step.
# Only metric extracts (CSV) come back; the .ncu-rep stays on the box.
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
bash tools/ncu_list.sh > gpurun_out/launches_step.txt
NL=$(grep -oE 'launches [0-9]+' gpurun_out/launches_step.txt | grep -oE '[0-9]+')
timeout 900 ncu --set full --clock-control none --import-source on -s $((3 * NL)) -c $NL -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_full_metrics.csv 2>/dev/null
ls -la gpurun_out/
rm -f gpurun_out/prof_final.ncu-rep
