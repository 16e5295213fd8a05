set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu2.log
timeout 600 python profiles/tune_chamfer.py 0 50 1 2 3 4 5 51 > gpurun_out/r02/tune_chamfer2.json 2> gpurun_out/r02/tune_chamfer2.err
tail -5 gpurun_out/r02/pytest_gpu2.log; cat gpurun_out/r02/tune_chamfer2.json
