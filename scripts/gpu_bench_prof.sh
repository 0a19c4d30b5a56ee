# Round bench + ncu evidence on one B200: plain bench first, then the launch list, then one full capture.
set -x
TAG=${1:-r01}
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_ref_$TAG.json
python bench.py --steps 2 --warmup 3 --batch 16384 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --batch 16384 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python bench.py --steps 1 --warmup 3 --batch 16384 --no-cpu-baseline --no-kernels > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:qp_solve -s 3 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 3 --batch 16384 --no-cpu-baseline --no-kernels > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
