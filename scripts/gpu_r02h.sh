set -x
TAG=${1:-r02h}
for W in 12 10; do
CMD2="python scripts/time_solver.py --kind trot --batch 16384 --steps 1 --oracle 0 --variants warp --env QPB_WARP_WARPS=$W"
ncu --set full --clock-control none --import-source on -k regex:qp_solve_warp -s 2 -c 1 -o gpurun_out/prof_warpv_${W}_$TAG -f $CMD2 > gpurun_out/ncu_full_${W}_$TAG.log 2>&1
done
ls -la gpurun_out | tail -5
