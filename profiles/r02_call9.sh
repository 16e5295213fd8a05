set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q -x -k "knn or group or Group or dgcnn or fullsize" > gpurun_out/r02/pytest_gpu9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu9.log
timeout 300 python profiles/tune_kernels.py knn > gpurun_out/r02/tune_knn9.log 2>&1
timeout 300 python profiles/configs_time.py > gpurun_out/r02/configs_time9.log 2>&1
tail -4 gpurun_out/r02/pytest_gpu9.log; cat gpurun_out/r02/tune_knn9.log; tail -25 gpurun_out/r02/configs_time9.log
