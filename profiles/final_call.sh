set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r02/bench_final3.json 2> gpurun_out/r02/bench_final3.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke.log 2>&1
for tool in racecheck synccheck initcheck; do timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_smoke.py > gpurun_out/r02/sanitize_$tool.log 2>&1; tail -2 gpurun_out/r02/sanitize_$tool.log; done
tail -3 gpurun_out/r02/pytest_gpu_final.log; cut -c1-300 gpurun_out/r02/bench_final3.json; cat gpurun_out/r02/smoke.log | tail -2
