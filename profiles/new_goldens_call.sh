mkdir -p gpurun_out/r03
timeout 38 python -m pytest tests/test_group_flavours.py tests/test_corrupt.py tests/test_ref_dgcnn.py -m gpu -q -k "group_classes or drop_patch or reference_module" 2>&1 | tail -n 15 | tee gpurun_out/r03/pytest_new_goldens.log
