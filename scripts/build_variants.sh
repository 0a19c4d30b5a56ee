# Build library variants of the generated warp kernels for one GPU sweep: each line of the list is
#   NAME ENV=VALUE ...   -> build/variants/NAME.so (generator switches: QPB_GEN_MAX_WARPS, QPB_GEN_PACK, QPB_GEN_VDIST, QPB_GEN_TDIST)
# usage: bash scripts/build_variants.sh "v10 QPB_GEN_PACK=0" "p12 QPB_GEN_MAX_WARPS=12" ...   (the default library is rebuilt last)
set -e
mkdir -p build/variants
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  env "$@" python -m apf_quadruped_b200.codegen_warp --out build/gen ${GEN_KINDS:-trot stance crawl} > /tmp/gen_$name.log 2>&1
  make -j8 apf_quadruped_b200/lib/libqpswift_b200.so > /tmp/build_$name.log 2>&1 || (grep -E "error" -A5 /tmp/build_$name.log | head -20; false)
  cp apf_quadruped_b200/lib/libqpswift_b200.so build/variants/$name.so
  echo "$name: $(grep -E 'registers|spill' build/obj/gen_warp_gen_trot.ptxas.log | grep -A1 'ELi1[0-9]ELb1' | tr '\n' ' ' | cut -c1-300)"
done
python -m apf_quadruped_b200.codegen_warp --out build/gen > /tmp/gen_default.log 2>&1
make -j8 apf_quadruped_b200/lib/libqpswift_b200.so > /tmp/build_default.log 2>&1
