set -x
mkdir -p gpurun_out/r02
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn3_kernel -c 1 -o gpurun_out/r02/knn3_group python profiles/tune_kernels.py knn > gpurun_out/r02/ncu_knn3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_reg_kernel -c 1 -o gpurun_out/r02/fps_reg python profiles/tune_kernels.py knn > gpurun_out/r02/ncu_fps.log 2>&1
ls -la gpurun_out/r02/*.ncu-rep
