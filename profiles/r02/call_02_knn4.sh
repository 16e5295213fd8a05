#!/bin/bash
# round 2, call 2: first run of knn4 (TMA tiles, multi-query warps): parity tests, sanitizer on a small case, A/B timing
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "knn or group or Group" -x 2>&1 | tail -15 > gpurun_out/r02/pytest_knn4_v5.txt
tail -5 gpurun_out/r02/pytest_knn4_v5.txt
timeout 600 python profiles/tune_knn.py --out gpurun_out/r02/tune_knn_v5.json > gpurun_out/r02/tune_knn_v5.log 2>&1
tail -12 gpurun_out/r02/tune_knn_v5.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r02/pytest_gpu_06.txt
tail -4 gpurun_out/r02/pytest_gpu_06.txt
