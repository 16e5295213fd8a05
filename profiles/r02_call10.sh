set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu10.log
compute-sanitizer --tool memcheck python profiles/sanitize_smoke.py > gpurun_out/r02/sanitize_memcheck.log 2>&1; tail -3 gpurun_out/r02/sanitize_memcheck.log
tail -4 gpurun_out/r02/pytest_gpu10.log
