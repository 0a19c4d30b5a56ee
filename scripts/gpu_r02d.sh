set -x
timeout 900 python -m pytest tests/test_rollout_gpu.py tests/test_wbc_gpu.py -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 3 --warmup 3 --config rollout --no-cpu-baseline --no-latency > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err; echo rc=$?; tail -c 600 gpurun_out/bench_r02d.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02d.json')); print(json.dumps(d['kernels'].get('wbc_formulate_kernel'), indent=1)); print(json.dumps(d['configs'], indent=1))"
