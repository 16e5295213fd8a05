set -x
mkdir -p gpurun_out/r02
timeout 600 python profiles/tune_chamfer.py 0 50 > gpurun_out/r02/tune_chamfer3.json 2> gpurun_out/r02/tune_chamfer3.err
PDAE_RECOVER_GROUPS=128 timeout 600 python profiles/tune_chamfer.py 0 > gpurun_out/r02/tune_chamfer3b.json 2>> gpurun_out/r02/tune_chamfer3.err
M=gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size,sm__cycles_elapsed.avg.per_second
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02/probe_chamfer.csv python profiles/probe_chamfer.py 0:128:2048 0:37:4096 0:74:4096 1:74:4096 1:37:4096 2:111:4096 2:37:4096 4:111:4096 100:37:4096 100:74:4096 > gpurun_out/r02/probe.log 2>&1
cat gpurun_out/r02/tune_chamfer3.json gpurun_out/r02/tune_chamfer3b.json
