set -x
mkdir -p gpurun_out/r02
timeout 600 python profiles/tune_chamfer.py 16 0 > gpurun_out/r02/tune_chamfer16.json 2> gpurun_out/r02/tune_chamfer16.err
cat gpurun_out/r02/tune_chamfer16.json; tail -3 gpurun_out/r02/tune_chamfer16.err
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu16.log
tail -6 gpurun_out/r02/pytest_gpu16.log
