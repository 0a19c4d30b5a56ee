set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 3 --warmup 3 --config trot65k --no-cpu-baseline --no-kernels > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err; echo rc=$?; tail -c 600 gpurun_out/bench_r02e.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02e.json')); print(json.dumps(d['latency'], indent=1)); print(d['value'], d['e2e'])"
QPB_NO_SINGLE_PATH=1 python bench.py --steps 3 --warmup 3 --config trot65k --no-cpu-baseline --no-kernels | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('old path', json.dumps(d['latency']['b200']))"
CMD="python bench.py --steps 1 --warmup 3 --config trot65k --no-cpu-baseline --no-latency"
ncu --set full --clock-control none --import-source on -k regex:wbc_formulate -s 2 -c 1 -o gpurun_out/prof_wbc_r02e -f $CMD > gpurun_out/ncu_wbc_r02e.log 2>&1
