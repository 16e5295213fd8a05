#!/bin/bash
mkdir -p gpurun_out/r02 /tmp/ncu
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-configs --no-ref-gpu --no-cpu-baseline > /tmp/ncu/launches.out 2>&1
# full captures of the round-2 kernels; only the summaries travel back (the reports exceed the 64 MiB return limit)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv1x1_tc|edge_stats|edge_forward|edge_backward|feat_dist_sym|feat_select2" -c 14 -o /tmp/ncu/row4 -f python profiles/ncu_row4.py > /tmp/ncu/row4.log 2>&1
python profiles/ncu_summary.py /tmp/ncu/row4.ncu-rep gpurun_out/r02/ncu_row4_featknn_summary.csv > gpurun_out/r02/ncu_row4_featknn_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chamfer_min_kernel|knn4_kernel|fps_reg" -s 30 -c 6 -o /tmp/ncu/step -f python bench.py --steps 2 --warmup 3 --no-configs --no-ref-gpu --no-cpu-baseline --no-graphs > /tmp/ncu/step.log 2>&1
python profiles/ncu_summary.py /tmp/ncu/step.ncu-rep gpurun_out/r02/ncu_step_kernels_summary.csv > gpurun_out/r02/ncu_step_kernels_summary.txt
grep -E "id|Kernel Name|time_duration|tensor|fma_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes" gpurun_out/r02/ncu_row4_featknn_summary.txt | head -80
