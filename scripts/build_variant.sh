# regenerate the warp kernels with the given environment switches and rebuild the library
set -e
make gen > /tmp/gen.log 2>&1
make -j8 apf_quadruped_b200/lib/libqpswift_b200.so > /tmp/buildv.log 2>&1 || (grep -E "error" -A5 /tmp/buildv.log | head -20; false)
grep -E "spill" build/obj/gen_warp_gen_trot.ptxas.log | sort | uniq -c | sort -rn | head -3
make ctest > /dev/null 2>&1
