export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned | cut -c 230-420; done
B200_ENGINE_INSTANCES=3 build/rest_replay --threads 3 --requests 60 --sizes 256 --pinned | cut -c 230-420
B200_ENGINE_PIPELINE_CHUNK=0 build/rest_replay --threads 2 --requests 60 --sizes 256 --pinned | cut -c 230-420
B200_ENGINE_PIPELINE_CHUNK=64 build/rest_replay --threads 2 --requests 60 --sizes 256 --pinned | cut -c 230-420
for t in 1 2 3; do build/rest_replay --threads $t --requests 60 --sizes 256 --pinned --uint8 | cut -c 230-420; done
B200_ENGINE_PIPELINE_CHUNK=0 build/rest_replay --threads 2 --requests 60 --sizes 256 --pinned --uint8 | cut -c 230-420
echo mixed
build/rest_replay --threads 32 --requests 1500 --pinned | cut -c 230-420
build/rest_replay --threads 32 --requests 1500 | cut -c 230-420
