# one GPU sweep over library variants (build/variants/*.so, scripts/build_variants.sh) x kernel switches
# usage: gpurun -- bash scripts/gpu_sweep.sh <tag> "<lib names>" "<time_solver variants>" [kind]
TAG=$1; LIBS=$2; VARS=$3; KIND=${4:-trot}
for L in $LIBS; do
  echo "== $L" >> gpurun_out/sweep_$TAG.txt
  QPB_LIB=$PWD/build/variants/$L.so timeout 600 python scripts/time_solver.py --kind $KIND --batch 65536 --steps 3 --oracle 64 --variants $VARS 2>&1 | grep -v "^+" | cut -c1-260 >> gpurun_out/sweep_$TAG.txt
done
cat gpurun_out/sweep_$TAG.txt
