#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn4_kernel -o gpurun_out/r02/ncu_knn4_v2 -f python profiles/ncu_knn.py H C4 > gpurun_out/r02/ncu_knn4_v2.log 2>&1
tail -3 gpurun_out/r02/ncu_knn4_v2.log
