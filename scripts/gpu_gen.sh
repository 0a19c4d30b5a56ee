# pattern-specialised (generated) throughput kernels against the interpreter: parity tests + timing
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_stream_gpu.py -x -q 2>&1 | tail -8
KINDS=trot,stance,crawl bash scripts/gpu_stream_variants.sh "X=1"
KINDS=trot COUNT=262144 bash scripts/gpu_stream_variants.sh "X=1"
