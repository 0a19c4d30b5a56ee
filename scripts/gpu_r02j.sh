set -x
TAG=${1:-r02j}
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_solver.py --kind stance --batch 65536 --steps 3 --oracle 64 --variants warp stream 2>&1 | grep -v "^+" | cut -c1-200
python scripts/time_solver.py --kind crawl --batch 65536 --steps 3 --oracle 64 --variants warp 2>&1 | grep -v "^+" | cut -c1-200
python scripts/time_solver.py --kind trot --batch 262144 --steps 3 --oracle 64 --variants warp stream --env QPB_STREAM_MIN=100000 2>&1 | grep -v "^+" | cut -c1-200
python scripts/time_solver.py --kind trot --batch 262144 --steps 3 --oracle 0 --variants warp 2>&1 | grep -v "^+" | cut -c1-200
