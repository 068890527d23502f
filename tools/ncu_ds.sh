mkdir -p gpurun_out /tmp/ncu
cap() {
  local P=$1 name=$2 regex=$3 skip=$4 TAG=$5
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$regex" --launch-skip $skip --launch-count 1 \
      -f -o /tmp/ncu/${P}_$name python tools/ncu_forward.py $P 256 2 > /tmp/ncu/${P}_$name.log 2>&1
  if [ -f /tmp/ncu/${P}_$name.ncu-rep ]; then
    python tools/ncu_summary.py /tmp/ncu/${P}_$name.ncu-rep 45 gpurun_out/${TAG}_sass_${P}_$name.tsv > gpurun_out/${TAG}_ncu_${P}_$name.txt 2>&1
  else
    echo "no report for $P $name"; tail -3 /tmp/ncu/${P}_$name.log
  fi
}
cap fp8 ds_b1 'dense_stream_kernel' 5 r02p
cap fp8 ds_b2 'dense_stream_kernel' 12 r02p
