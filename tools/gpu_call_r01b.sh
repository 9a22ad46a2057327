mkdir -p gpurun_out
( timeout 135 python -m pytest tests/test_gpu_crs_utils.py tests/test_gpu_spmm.py tests/test_shim.py -x -q -m "gpu or gpu_next" -k "not test_spmm_sweep or 1000-3-200-10" -p no:cacheprovider > gpurun_out/pytest_r01b.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r01b.log ) &
./kokkos-kernels_b200/lib/gpu_check --suite spmm --spmm-scale 23 --out gpurun_out/gpu_check_r01b.jsonl > gpurun_out/gpu_check_r01b.log 2>&1
./kokkos-kernels_b200/lib/gpu_check --big --suite spmm_sweep --suite crs_big --out gpurun_out/gpu_check_r01b.jsonl >> gpurun_out/gpu_check_r01b.log 2>&1
./kokkos-kernels_b200/lib/shim_driver > gpurun_out/shim_r01b.log 2>&1; echo "shim exit=$?" >> gpurun_out/shim_r01b.log
timeout 60 ncu --set full --import-source on --clock-control none --target-processes all -k regex:spmm_tile_kernel -c 1 -f -o gpurun_out/r01_spmm_tile ./kokkos-kernels_b200/lib/gpu_check --big --suite spmm --out gpurun_out/gpu_check_ncu_scratch.jsonl > gpurun_out/ncu_r01b.log 2>&1
wait
grep -E "FAIL|summary" gpurun_out/gpu_check_r01b.log | head -30; tail -3 gpurun_out/shim_r01b.log; tail -5 gpurun_out/pytest_r01b.log; tail -3 gpurun_out/ncu_r01b.log; ls -la gpurun_out
