timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -25
for p in bf16 fp8; do timeout 300 python tools/gpu_probe.py $p 256 10 --check --profile --e2e 2>&1 | tail -16; done
timeout 120 python tools/gpu_probe.py bf16 1 50 2>&1 | tail -2
timeout 120 python tools/gpu_probe.py fp32 1 50 2>&1 | tail -2
