set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_inst2.json 2> gpurun_out/bench_inst2.err; tail -c 600 gpurun_out/bench_inst2.err
export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0
for mode in "" "--pinned" "--uint8" "--uint8 --pinned"; do
  B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 32 --requests 1500 $mode
done
for mode in "" "--pinned"; do
  B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 64 --requests 20000 --sizes 1 $mode
  B200_ENGINE_COALESCE_US=0 build/rest_replay --threads 64 --requests 5000 --sizes 1 $mode
done
B200_ENGINE_INSTANCES=1 build/rest_replay --threads 32 --requests 1500 --pinned
