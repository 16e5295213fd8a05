set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu_final2.log
timeout 900 python bench.py > gpurun_out/r02/bench_final4.json 2> gpurun_out/r02/bench_final4.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke2.log 2>&1
timeout 600 python profiles/configs_time.py > gpurun_out/r02/configs_final.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_smoke.py > gpurun_out/r02/sanitize_memcheck2.log 2>&1; tail -2 gpurun_out/r02/sanitize_memcheck2.log
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitize_smoke.py > gpurun_out/r02/sanitize_racecheck2.log 2>&1; tail -2 gpurun_out/r02/sanitize_racecheck2.log
tail -3 gpurun_out/r02/pytest_gpu_final2.log; cut -c1-300 gpurun_out/r02/bench_final4.json; tail -1 gpurun_out/r02/smoke2.log; tail -24 gpurun_out/r02/configs_final.log
