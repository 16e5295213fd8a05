set -x
mkdir -p gpurun_out/r02
timeout 900 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02/bench_ref.json 2> gpurun_out/r02/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_bench_final.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-ref-gpu --no-graphs > gpurun_out/r02/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chamfer_min_kernel|chamfer_col_recover|knn3_kernel|fps_reg|chamfer_bwd|chamfer_loss" -c 9 -o gpurun_out/r02/final_kernels python profiles/probe_timeline.py > gpurun_out/r02/ncu_final.log 2>&1
cut -c1-400 gpurun_out/r02/bench_final.json; cut -c1-300 gpurun_out/r02/bench_ref.json; tail -2 gpurun_out/r02/bench_final.err
