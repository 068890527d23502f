timeout 900 python -m pytest tests/ -q -m gpu 2>&1 | tail -15
for p in bf16 fp8; do timeout 300 python tools/gpu_probe.py $p 256 10 --check --profile 2>&1 | tail -12; done
timeout 120 python tools/gpu_probe.py bf16 1 50 2>&1 | tail -1
