python scripts/time_solver.py --kind trot --batch 65536 --steps 3 --oracle 64 --variants warp $EXTRA 2>&1 | grep -v "^+" | cut -c1-200
