set -x
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_wbc_gpu.py -m gpu -q 2>&1 | tail -25
python bench.py --steps 3 --warmup 3 --config trot65k --no-cpu-baseline --no-latency > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo rc=$?; tail -c 400 gpurun_out/bench_r02c.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02c.json')); print(json.dumps(d['kernels'], indent=1)); print(d['value'], d['e2e'])"
