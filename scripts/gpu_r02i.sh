set -x
CMD2="python scripts/time_solver.py --kind trot --batch 16384 --steps 1 --oracle 0 --variants warp"
ncu --set full --clock-control none --import-source on -k regex:qp_solve_warp -s 2 -c 1 -o gpurun_out/prof_warp_r02i -f $CMD2 > gpurun_out/ncu_full_r02i.log 2>&1
