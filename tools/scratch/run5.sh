timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "conv3x3 or dense_block" 2>&1 | tail -15
for p in bf16 fp8; do timeout 300 python tools/gpu_probe.py $p 256 10 --check --profile 2>&1 | tail -9; done
