#!/bin/bash
for cfg in "0 0" "64 0" "64 16" "200 8"; do
  set -- $cfg
  echo "== sleep $1 backoff_after $2"
  B200_EXTRA_NVCC="-DB200_POLL_SLEEP_NS=$1 -DB200_POLL_BACKOFF_AFTER=$2" python build_engine.py --force > /dev/null 2>&1
  timeout 200 python tools/gpu_probe.py fp8 256 10 --profile 2>&1 | grep -E "fp8 bs256"
  python - <<EOF
import json
d=json.load(open("gpurun_out/steps_fp8_bs256.json"))
print([ (x["name"][-28:-10], round(x["ms"]*1e3)) for x in d if x["ms"]>0.15])
EOF
done
python build_engine.py --force > /dev/null 2>&1
