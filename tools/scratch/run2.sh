for c in conv1x1_bn_relu conv1x1_partial_chunk conv3x3 cout256 transition stem_maxpool dense_block gap_gemm_softmax; do
  echo "=== $c bf16"; timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_operator_graphs_match_oracle and $c and bf16" -x 2>&1 | tail -6
done
echo "=== fp32 all ops"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "test_operator_graphs_match_oracle and fp32" 2>&1 | tail -6
