timeout 200 python tools/gpu_probe.py bf16 256 1 > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 0 -c 1 -f -o gpurun_out/prof_halo_bf16 python tools/gpu_probe.py bf16 256 1 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 3 -c 1 -f -o gpurun_out/prof_umma_bf16 python tools/gpu_probe.py bf16 256 1 > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu1.log gpurun_out/ncu2.log
