set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu14.log
timeout 600 python profiles/tune_chamfer.py 0 25 50 > gpurun_out/r02/tune_chamfer14.json 2> gpurun_out/r02/tune_chamfer14.err
tail -4 gpurun_out/r02/pytest_gpu14.log; cat gpurun_out/r02/tune_chamfer14.json
