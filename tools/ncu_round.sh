#!/bin/bash
# usage: tools/ncu_round.sh <tag>      (run under gpurun, ONE GPU; text summaries -> gpurun_out/, copy what should be judged to profiles/)
# One `ncu --set full` capture per kernel family of the final tree, on the bs256 forward of each precision mode (CUDA graphs off so
# that every kernel is a plain launch).  name:regex:launches-to-skip; the regex is matched against the demangled kernel name.
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
cap() {  # precision name regex skip
  local P=$1 name=$2 regex=$3 skip=$4
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" --launch-skip $skip --launch-count 1 \
      -f -o /tmp/ncu/${P}_$name python tools/ncu_forward.py $P 256 2 > /tmp/ncu/${P}_$name.log 2>&1
  if [ -f /tmp/ncu/${P}_$name.ncu-rep ]; then
    rm -f gpurun_out/${TAG}_sass_${P}_$name.tsv
    python tools/ncu_summary.py /tmp/ncu/${P}_$name.ncu-rep 30 gpurun_out/${TAG}_sass_${P}_$name.tsv > gpurun_out/${TAG}_ncu_${P}_$name.txt 2>&1
  else
    echo "no report for $P $name"; tail -3 /tmp/ncu/${P}_$name.log
  fi
}
for P in fp8 bf16 fp32; do
  python tools/ncu_forward.py $P 256 2 > gpurun_out/ncu_plain_$P.log 2>&1 || { echo "plain run failed for $P"; tail -5 gpurun_out/ncu_plain_$P.log; exit 1; }
done
cap fp8 c1x1_cin224 'conv1x1_tma_kernel' 5
cap fp8 c3x3_b1 'conv3x3_tma_kernel' 5
cap fp8 c3x3_b2 'conv3x3_tma_kernel' 17
cap fp8 dense_b3 'dense_block_kernel' 0
cap fp8 dense_b4 'dense_block_kernel' 1
cap fp8 stem 'stem_conv7x7_kernel' 0
cap fp8 maxpool 'maxpool3x3s2_kernel' 0
cap fp8 poolbn_t1 'pool_bn_relu_2x2_kernel' 0
cap fp8 gap 'gap_kernel' 0
cap fp8 fc 'fc_f32' 0
cap bf16 c1x1_cin224 'conv1x1_tma_kernel' 5
cap bf16 c3x3_b1 'conv3x3_tma_kernel' 5
cap bf16 stem 'stem_conv7x7_kernel' 0
cap fp32 c1x1_cin224 'conv1x1_f32x3_kernel' 5
cap fp32 c1x1_b3 'conv1x1_f32x3_kernel' 30
cap fp32 c3x3_b1 'conv3x3_f32x3_kernel' 5
cap fp32 stem 'stem_conv7x7_kernel' 0
cap fp32 poolbn_t1 'pool_bn_relu_2x2_f32_kernel' 0
ls gpurun_out | grep ${TAG}_ncu | wc -l
