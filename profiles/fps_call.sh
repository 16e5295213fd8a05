set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -k "fps or FPS or golden or fullsize" > gpurun_out/r02/pytest_fps2.log 2>&1; tail -2 gpurun_out/r02/pytest_fps2.log
timeout 600 python profiles/tune_kernels.py fps > gpurun_out/r02/tune_fps2.log 2>&1; grep -E "8192|4096|None" gpurun_out/r02/tune_fps2.log
timeout 300 python profiles/configs_time.py 2>&1 | grep -E "fps" 
