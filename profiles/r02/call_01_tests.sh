#!/bin/bash
# round 2, call 1: GPU-made pointnet2 golden + the whole GPU suite (no -x) + a bench line
mkdir -p gpurun_out/r02
python tests/golden/make_golden_pointnet2.py --gpu > gpurun_out/r02/golden_pointnet2_gpu.log 2>&1
cp gpurun_out/pointnet2_ref_gpu.npz tests/golden/pointnet2_ref_gpu.npz
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02/pytest_gpu_01.txt
tail -5 gpurun_out/r02/pytest_gpu_01.txt
timeout 600 python bench.py > gpurun_out/r02/bench_01.json 2> gpurun_out/r02/bench_01.err
tail -c 600 gpurun_out/r02/bench_01.json
