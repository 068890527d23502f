set -u
export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=all B200_ENGINE_COALESCE_US=0
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" 2>&1 | tail -2
{
build/rest_replay --threads 64 --requests 12000 --pinned
build/rest_replay --threads 64 --requests 12000
build/rest_replay --threads 64 --requests 12000 --pinned --uint8
build/rest_replay --threads 128 --requests 12000 --pinned
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 256 --requests 60000 --sizes 1 --pinned
B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 256 --requests 30000 --sizes 1 --pinned
} > gpurun_out/r01k_replay_8gpu.jsonl 2>&1
cut -c 200-460 gpurun_out/r01k_replay_8gpu.jsonl
unset B200_ENGINE_PRECISION B200_ENGINE_DEVICES B200_ENGINE_COALESCE_US
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r01k_bench_fp8_8gpu.json 2> gpurun_out/r01k_bench_fp8_8gpu.err
tail -c 1500 gpurun_out/r01k_bench_fp8_8gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['e2e'], d['e2e_uint8'])" || tail -5 gpurun_out/r01k_bench_fp8_8gpu.err
