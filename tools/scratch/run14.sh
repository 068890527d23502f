timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_kernel or instances" 2>&1 | tail -4
bash tools/run_ncu2.sh fp8 r01k dense_b4:dense_block_kernel:3 > /dev/null 2>&1
B200_ENGINE_TILEFUSE=1 bash tools/run_ncu2.sh fp8 r01k tile_b1_cin224:dense_tile_kernel:16 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01k_launches_fp8.csv python tools/ncu_forward.py fp8 256 2 > /dev/null 2>&1
head -12 gpurun_out/r01k_ncu_fp8_dense_b4.txt; head -12 gpurun_out/r01k_ncu_fp8_tile_b1_cin224.txt; wc -l gpurun_out/r01k_launches_fp8.csv
