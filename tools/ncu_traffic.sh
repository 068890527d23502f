#!/bin/bash
# usage: tools/ncu_traffic.sh <precision> <batch> <tag>     (run under gpurun, one GPU)
# DRAM traffic of ONE forward: dram__bytes_read.sum + dram__bytes_write.sum of every launch, graphs off so each kernel is a plain
# launch.  Writes profiles-ready files to gpurun_out/: <tag>_traffic_<precision>_bs<batch>.csv (+ .json with the forward count),
# which bench.py's roofline.traffic reads once they are copied to profiles/.
P=$1; B=$2; TAG=$3
mkdir -p gpurun_out
python tools/ncu_forward.py $P $B 1 > gpurun_out/ncu_plain_${P}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_${P}.log; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_traffic_${P}_bs${B}.csv python tools/ncu_forward.py $P $B 1 > gpurun_out/ncu_traffic_${P}.log 2>&1
echo "{\"forwards\": 1, \"precision\": \"$P\", \"batch\": $B, \"command\": \"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none python tools/ncu_forward.py $P $B 1\"}" > gpurun_out/${TAG}_traffic_${P}_bs${B}.json
tail -3 gpurun_out/${TAG}_traffic_${P}_bs${B}.csv
