# round 2, first GPU pass: tests, kernel variants (warps per SM sweep of the warp-per-QP kernel), bench, ncu full of the warp kernel
set -x
TAG=${1:-r02a}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_solver.py --kind trot --batch 65536 --steps 3 --variants stream warp warp_free warp7 warp6 warp4 warp2 > gpurun_out/variants_trot_$TAG.txt 2>&1
python scripts/time_solver.py --kind stance --batch 65536 --steps 3 --oracle 64 --variants stream warp warp4 > gpurun_out/variants_stance_$TAG.txt 2>&1
python scripts/time_solver.py --kind crawl --batch 65536 --steps 3 --oracle 64 --variants stream warp > gpurun_out/variants_crawl_$TAG.txt 2>&1
cat gpurun_out/variants_*_$TAG.txt | grep -v "^+"
python bench.py --steps 5 --warmup 3 --no-kernels > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_$TAG.json
CMD2="python scripts/time_solver.py --kind trot --batch 16384 --steps 1 --oracle 0 --variants warp"
ncu --set full --clock-control none --import-source on -k regex:qp_solve_warp -s 2 -c 1 -o gpurun_out/prof_warp_$TAG -f $CMD2 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
