export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
c() { cut -c 230-420; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo mixed
for i in 4 6 8; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 4000 --pinned | c; done
for i in 4 8; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 8000 | c; done
for i in 4 8; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 8000 --uint8 | c; done
echo bs1
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 64 --requests 20000 --sizes 1  | c
B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 64 --requests 20000 --sizes 1  | c
B200_ENGINE_COALESCE_US=0 B200_ENGINE_INSTANCES=8 build/rest_replay --threads 64 --requests 20000 --sizes 1  | c
unset B200_ENGINE_PRECISION B200_ENGINE_DEVICES B200_ENGINE_COALESCE_US
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_inst4.json 2> gpurun_out/bench_inst4.err; tail -c 300 gpurun_out/bench_inst4.err
