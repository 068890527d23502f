timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batch_properties or concurrent or validation" 2>&1 | tail -3
for c in 0 32 64 128; do echo "== chunk $c"; B200_ENGINE_PIPELINE_CHUNK=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']), 'lat', d['latency'])"; done
for b in 32 64 128; do timeout 100 python tools/gpu_probe.py fp8 $b 20 | tail -1; done
