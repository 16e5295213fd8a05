set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02/pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu7.log
timeout 600 python bench.py --steps 300 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02/bench_fused.json 2> gpurun_out/r02/bench_fused.err
timeout 300 python profiles/tune_kernels.py knn > gpurun_out/r02/tune_knn7.log 2>&1
tail -4 gpurun_out/r02/pytest_gpu7.log; cut -c1-330 gpurun_out/r02/bench_fused.json; tail -3 gpurun_out/r02/bench_fused.err; cat gpurun_out/r02/tune_knn7.log
