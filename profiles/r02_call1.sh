set -x
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/smi.txt
python tests/golden/make_golden_interp.py gpurun_out/golden > gpurun_out/r02/golden_interp.log 2>&1
cp gpurun_out/golden/interp.npz tests/golden/interp.npz
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02/pytest_gpu.log
./profiles/ubench/issue_mix > gpurun_out/r02/ubench_issue_mix.json 2> gpurun_out/r02/ubench_issue_mix.err
timeout 600 python profiles/tune_chamfer.py > gpurun_out/r02/tune_chamfer.json 2> gpurun_out/r02/tune_chamfer.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chamfer_min_kernel -c 1 -o gpurun_out/r02/chamfer_sym_v0 python profiles/tune_chamfer.py 0 > gpurun_out/r02/ncu_v0.log 2>&1
timeout 600 python bench.py > gpurun_out/r02/bench_a.json 2> gpurun_out/r02/bench_a.err
tail -3 gpurun_out/r02/pytest_gpu.log; cat gpurun_out/r02/tune_chamfer.json; cat gpurun_out/r02/ubench_issue_mix.json; cat gpurun_out/r02/bench_a.json | cut -c1-400
