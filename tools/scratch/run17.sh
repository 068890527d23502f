export B200_ENGINE_PRECISION=fp8 B200_ENGINE_DEVICES=0 B200_ENGINE_COALESCE_US=0
for i in 1 4 8; do s=$(date +%s.%N); B200_ENGINE_INSTANCES=$i build/rest_replay --threads 32 --requests 4000 --pinned > /tmp/o.json; e=$(date +%s.%N); python - $i $s $e <<'P'
import json,sys
d=json.load(open('/tmp/o.json')); print('instances',sys.argv[1],'wall',round(float(sys.argv[3])-float(sys.argv[2]),2),'s img/s',d['images_per_s'],'p50',d['latency_ms_p50'],'p99',d['latency_ms_p99'])
P
done
