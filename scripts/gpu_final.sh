# round-end evidence: GPU tests, smoke, bench (both arms), ncu launch list and full capture of the bench command
set -x
TAG=${1:-r01c}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernels"
$CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-qp_solve_stream} -s 3 -c 1 -o gpurun_out/prof_$TAG -f $CMD2 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -c 600 gpurun_out/bench_$TAG.json
