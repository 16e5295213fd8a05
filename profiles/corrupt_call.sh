set -x
mkdir -p gpurun_out/r03
timeout 240 python -m pytest tests/test_corrupt.py -m gpu -q > gpurun_out/r03/pytest_corrupt.log 2>&1; tail -4 gpurun_out/r03/pytest_corrupt.log
timeout 120 python profiles/corrupt_time.py > gpurun_out/r03/corrupt_time.json 2> gpurun_out/r03/corrupt_time.err; cat gpurun_out/r03/corrupt_time.json; tail -2 gpurun_out/r03/corrupt_time.err
timeout 300 python bench.py > gpurun_out/r03/bench.json 2> gpurun_out/r03/bench.err; cut -c1-700 gpurun_out/r03/bench.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03/smoke.log 2>&1; tail -1 gpurun_out/r03/smoke.log
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_corrupt.py > gpurun_out/r03/pytest_gpu.log 2>&1; tail -4 gpurun_out/r03/pytest_gpu.log
