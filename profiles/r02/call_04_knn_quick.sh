#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python profiles/tune_knn.py --quick --out gpurun_out/r02/tune_knn_v6_auto.json 2>&1 | grep -v "^{" | tail -6
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
