set -x
TAG=${1:-r02g}
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_stream_gpu.py tests/test_boundary_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -5
python scripts/time_solver.py --kind trot --batch 65536 --steps 3 --variants warp warp11 warp10 warp9 warp8 warp6 > gpurun_out/variants_trot_$TAG.txt 2>&1
python scripts/time_solver.py --kind stance --batch 65536 --steps 3 --oracle 64 --variants warp warp8 > gpurun_out/variants_stance_$TAG.txt 2>&1
python scripts/time_solver.py --kind crawl --batch 65536 --steps 3 --oracle 64 --variants warp warp8 > gpurun_out/variants_crawl_$TAG.txt 2>&1
cat gpurun_out/variants_*_$TAG.txt | grep -v "^+" | cut -c1-260
