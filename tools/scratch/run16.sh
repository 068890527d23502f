export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
c() { cut -c 230-420; }
for i in 1 4 8; do /usr/bin/time -f "load+run wall %es" env B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 4000 --pinned 2>&1 | c; done
B200_ENGINE_INSTANCES=4 build/rest_replay --threads 32 --requests 6000 | c
