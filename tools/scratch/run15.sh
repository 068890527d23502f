export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
c() { cut -c 300-400; }
for ch in 128 64 32 0; do echo chunk $ch; B200_ENGINE_PIPELINE_CHUNK=$ch build/rest_replay --threads 1 --requests 60 --sizes 256 --pinned | c; B200_ENGINE_PIPELINE_CHUNK=$ch build/rest_replay --threads 1 --requests 60 --sizes 256 --pinned --uint8 | c; B200_ENGINE_PIPELINE_CHUNK=$ch build/rest_replay --threads 1 --requests 100 --sizes 128 --pinned | c; B200_ENGINE_PIPELINE_CHUNK=$ch build/rest_replay --threads 1 --requests 100 --sizes 64 --pinned | c; done
