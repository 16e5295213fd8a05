set -x
mkdir -p gpurun_out/r02
PDAE_RECOVER_GROUPS=128 timeout 600 python profiles/tune_chamfer.py 0 25 50 > gpurun_out/r02/tune_chamfer15.json 2> gpurun_out/r02/tune_chamfer15.err
cat gpurun_out/r02/tune_chamfer15.json
PDAE_RECOVER_GROUPS=128 timeout 600 python -m pytest tests -m gpu -q -k "chamfer or fullsize" > gpurun_out/r02/pytest_gpu15.log 2>&1; tail -3 gpurun_out/r02/pytest_gpu15.log
