#!/bin/bash
# Round measurements on one B200 (run under gpurun): GPU tests, the bench line of every BASELINE.json single-GPU config, and
# the mixed-batch replay (configs[4]) through the C-ABI.  Outputs land in gpurun_out/ (copy what should be judged to profiles/).
set -u
TAG=${1:-r01k}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_fp8_bs256.json 2> gpurun_out/${TAG}_bench_fp8_bs256.err
python bench.py --steps 20 --warmup 5 --precision bf16 --batch 64 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16_bs64.json 2> gpurun_out/${TAG}_bench_bf16_bs64.err
python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16_bs256.json 2> gpurun_out/${TAG}_bench_bf16_bs256.err
python bench.py --steps 10 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32_bs256.json 2> gpurun_out/${TAG}_bench_fp32_bs256.err
export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=${REPLAY_DEVICES:-0}
{
for mode in "--pinned" "" "--pinned --uint8" "--uint8"; do
  B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 32 --requests 6000 $mode
done
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 64 --requests 30000 --sizes 1 --pinned
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 64 --requests 30000 --sizes 1
B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 64 --requests 10000 --sizes 1 --pinned
} > gpurun_out/${TAG}_replay_1gpu.jsonl 2>&1
tail -n 3 gpurun_out/${TAG}_pytest_gpu.txt
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), 'img/s', round(d['ms_per_step'],3),'ms | e2e',round(d['e2e']['value']), 'serial', round(d['e2e'].get('serial_value',0)), '| u8', round(d['e2e_uint8']['value']), '| lat', d.get('latency'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
P
done
cut -c 200-460 gpurun_out/${TAG}_replay_1gpu.jsonl
