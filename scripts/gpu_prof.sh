# ncu full capture of the solve kernel (+ plain run first).  usage: gpu_prof.sh TAG KIND BATCH [ENV...]
set -x
TAG=${1:-x}
KIND=${2:-trot}
BATCH=${3:-65536}
shift 3
CMD="python bench.py --steps 1 --warmup 3 --batch $BATCH --kind $KIND --no-cpu-baseline --no-kernels"
env "$@" $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
env "$@" ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-qp_solve_stream} -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log
