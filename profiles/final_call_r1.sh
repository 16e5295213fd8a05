mkdir -p gpurun_out/r03
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r03/pytest_gpu_final.log 2>&1; tail -n 3 gpurun_out/r03/pytest_gpu_final.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03/smoke_final.log 2>&1; tail -n 1 gpurun_out/r03/smoke_final.log
timeout 200 python bench.py > gpurun_out/r03/bench_final_r1.json 2> gpurun_out/r03/bench_final_r1.err; cut -c1-330 gpurun_out/r03/bench_final_r1.json
