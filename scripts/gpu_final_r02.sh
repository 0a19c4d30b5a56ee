# round-2 evidence on one B200: GPU tests, smoke, the whole bench line, the reference arm, launch list, full capture of the dominant kernel at the headline size
set -x
TAG=${1:-r02k}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
CMD="python bench.py --steps 2 --warmup 3 --config trot65k --no-cpu-baseline --no-latency"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
CMD2="python scripts/time_solver.py --kind trot --batch 65536 --steps 1 --oracle 0 --variants warp"
$CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_solve_warp -s 2 -c 1 -o gpurun_out/prof_warp65k_$TAG -f $CMD2 > gpurun_out/ncu_full_$TAG.log 2>&1
for K in stance crawl; do
CMD3="python scripts/time_solver.py --kind $K --batch 65536 --steps 1 --oracle 0 --variants warp"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:qp_solve_warp -s 2 -c 1 --csv --log-file gpurun_out/traffic_${K}_$TAG.csv $CMD3 > /dev/null 2>&1
done
ls -la gpurun_out | tail -12
