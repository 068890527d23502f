nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_gpu" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2_err.log | tail -1 > gpurun_out/bench_fp8_2gpu.json; tail -3 gpurun_out/bench2_err.log; cat gpurun_out/bench_fp8_2gpu.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/bench1_err.log | tail -1 > gpurun_out/bench_fp8_1gpu.json; tail -3 gpurun_out/bench1_err.log; cat gpurun_out/bench_fp8_1gpu.json
timeout 600 python bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
