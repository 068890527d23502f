export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
c() { cut -c 230-420; }
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned | c; done
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned --uint8 | c; done
echo mixed
for i in 2 4; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 3000 --pinned | c; done
B200_ENGINE_CHAIN_MIN_BATCH=32 build/rest_replay --threads 32 --requests 3000 --pinned | c
B200_ENGINE_CHAIN_MIN_BATCH=128 build/rest_replay --threads 32 --requests 3000 --pinned | c
build/rest_replay --threads 32 --requests 3000 --pinned --uint8 | c
echo pageable
build/rest_replay --threads 32 --requests 8000 | c
B200_ENGINE_STAGE_PAGEABLE=0 build/rest_replay --threads 32 --requests 3000 | c
build/rest_replay --threads 8 --requests 8000 | c
B200_ENGINE_STAGE_PAGEABLE=0 build/rest_replay --threads 8 --requests 3000 | c
build/rest_replay --threads 32 --requests 8000 --uint8 | c
B200_ENGINE_STAGE_PAGEABLE=0 build/rest_replay --threads 32 --requests 3000 --uint8 | c
nproc; grep -m1 "model name" /proc/cpuinfo
