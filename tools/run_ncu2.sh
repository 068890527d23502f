#!/bin/bash
# usage: tools/run_ncu2.sh <precision> <tag> name:regex:skip ...   (run under gpurun; text summaries -> gpurun_out/)
P=$1; TAG=$2; shift 2
mkdir -p gpurun_out /tmp/ncu
python tools/ncu_forward.py $P 256 2 > gpurun_out/ncu_plain_$P.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$P.log; exit 1; }
for spec in "$@"; do
  IFS=: read name regex skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k "regex:$regex" --launch-skip $skip --launch-count 1 -f -o /tmp/ncu/$name python tools/ncu_forward.py $P 256 2 > /tmp/ncu/$name.log 2>&1
  rm -f gpurun_out/${TAG}_sass_${P}_$name.tsv; python tools/ncu_summary.py /tmp/ncu/$name.ncu-rep 40 gpurun_out/${TAG}_sass_${P}_$name.tsv > gpurun_out/${TAG}_ncu_${P}_$name.txt 2>&1
done
ls gpurun_out
