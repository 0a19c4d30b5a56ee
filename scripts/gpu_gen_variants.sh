# library variants (apf_quadruped_b200/lib/variants/*.so, e.g. generated kernels with different chunk sizes): parity + timing
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in apf_quadruped_b200/lib/variants/*.so; do
  echo "##### $(basename $lib)"
  QPB_LIB=$lib timeout 600 python -m pytest tests/test_stream_gpu.py -x -q 2>&1 | tail -2
  for cnt in ${COUNTS:-65536 262144}; do
    QPB_LIB=$lib COUNT=$cnt KINDS=${KINDS:-trot} bash scripts/gpu_stream_variants.sh "X=1" | grep RESULT
  done
done
