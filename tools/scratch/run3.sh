export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
c() { cut -c 230-420; }
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned | c; done
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned --uint8 | c; done
B200_ENGINE_INSTANCES=3 build/rest_replay --threads 3 --requests 60 --sizes 256 --pinned --uint8 | c
B200_ENGINE_PIPELINE_CHUNK=0 build/rest_replay --threads 2 --requests 60 --sizes 256 --pinned --uint8 | c
echo pageable 256
for t in 1 2 4; do build/rest_replay --threads $t --requests 40 --sizes 256 | c; done
B200_ENGINE_STAGE_PAGEABLE=0 build/rest_replay --threads 2 --requests 40 --sizes 256 | c
echo mixed
for i in 2 3 4; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 1500 --pinned | c; done
for i in 2 4; do B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 1500 | c; done
B200_ENGINE_INSTANCES=4 build/rest_replay --threads 32 --requests 1500 --uint8 | c
B200_ENGINE_CHAIN=0 build/rest_replay --threads 32 --requests 1500 --pinned | c
echo bs1
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 64 --requests 20000 --sizes 1  | c
B200_ENGINE_COALESCE_US=200 build/rest_replay --threads 256 --requests 40000 --sizes 1  | c
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
