timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "operator_graphs or low_precision" 2>&1 | tail -5
for p in bf16 fp8; do timeout 300 python tools/gpu_probe.py $p 256 10 --check --profile 2>&1 | tail -9; done
timeout 120 python tools/gpu_probe.py bf16 1 50 2>&1 | tail -1
