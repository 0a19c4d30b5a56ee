set -x
timeout 600 python -m pytest tests/test_wbc_gpu.py tests/test_rollout_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --config trot65k --no-cpu-baseline --no-latency | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['kernels']['wbc_formulate_kernel']))"
