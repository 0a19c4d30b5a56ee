# ncu full capture of the solve kernel on a small trot batch + source page CSV
set -x
TAG=${1:-x}
KIND=${2:-trot}
python bench.py --steps 1 --warmup 3 --batch 16384 --kind $KIND --no-cpu-baseline --no-kernels > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_solve -s 3 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch 16384 --kind $KIND --no-cpu-baseline --no-kernels > gpurun_out/ncu_full_$TAG.log 2>&1
cat gpurun_out/plain_$TAG.log | tail -2
