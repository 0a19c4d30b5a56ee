# library variants (apf_quadruped_b200/lib/variants/*.so, e.g. generated kernels with different options): timing via QPB_LIB
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in apf_quadruped_b200/lib/variants/*.so; do
  echo "##### $(basename $lib)"
  for cnt in ${COUNTS:-65536 262144}; do
    QPB_LIB=$lib COUNT=$cnt KINDS=${KINDS:-trot} bash scripts/gpu_stream_variants.sh "X=1" | grep RESULT
  done
done
