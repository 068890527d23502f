#!/bin/bash
# usage: tools/run_ncu.sh <precision> <tag>   (run under gpurun; writes text summaries to gpurun_out/)
P=${1:-fp8}; TAG=${2:-r01}
mkdir -p gpurun_out /tmp/ncu
python tools/ncu_forward.py $P 256 2 > gpurun_out/ncu_plain_$P.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$P.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$P.csv python tools/ncu_forward.py $P 256 2 > /tmp/ncu/ll.log 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f -o /tmp/ncu/$1 python tools/ncu_forward.py $P 256 2 > /tmp/ncu/$1.log 2>&1
  python tools/ncu_summary.py /tmp/ncu/$1.ncu-rep 28 > gpurun_out/${TAG}_ncu_${P}_$1.txt 2>&1
}
cap stem stem_conv7x7 1
cap c1x1_cin64 conv_umma 61
cap c1x1_cin224 conv_umma 66
cap trans1 conv_umma 67
cap c1x1_b3_cin640 conv_umma 93
cap halo_b1 conv3x3_halo 58
cap halo_b3 conv3x3_halo 88
cap maxpool pool_kernel 1
ls -la /tmp/ncu gpurun_out | head -40
