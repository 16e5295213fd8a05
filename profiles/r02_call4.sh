set -x
mkdir -p gpurun_out/r02
timeout 600 python profiles/tune_chamfer.py 0 12 13 6 > gpurun_out/r02/tune_chamfer4.json 2> gpurun_out/r02/tune_chamfer4.err
M=gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size,sm__cycles_elapsed.avg.per_second,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg
timeout 900 ncu --metrics $M --clock-control none -k regex:chamfer_min --csv --log-file gpurun_out/r02/probe_chamfer4.csv python profiles/probe_chamfer.py 0:128:2048 1:128:2048 2:128:2048 5:128:2048 6:128:2048 7:128:2048 12:128:2048 13:128:2048 > gpurun_out/r02/probe4.log 2>&1
cat gpurun_out/r02/tune_chamfer4.json
