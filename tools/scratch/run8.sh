timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/bench_fp8.json; tail -5 gpurun_out/bench_err.log; cat gpurun_out/bench_fp8.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1
nvidia-smi -L
