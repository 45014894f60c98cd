set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 228 -c 76 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"conv_umma|lateral0|roi_align|head_attention|pool_partial|decode_kernel|stem|head_final" -s 33 -c 11 -o gpurun_out/prof_final_a -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
ncu -i gpurun_out/prof_final_a.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_final_a.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"pw_mma|dw_tiled|se_fc" -s 168 -c 56 -o gpurun_out/prof_final_b -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
ncu -i gpurun_out/prof_final_b.ncu-rep --page raw --csv --metrics $M > gpurun_out/ncu_final_b.csv 2>/dev/null
ls -la gpurun_out/
rm -f gpurun_out/prof_final_b.ncu-rep
