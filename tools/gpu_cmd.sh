python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_configs.py --scale 0.5 --only 5 2>&1 | tail -1 | cut -c1-230
GL_DISABLE_TILE=1 python tools/bench_configs.py --scale 0.3 --only 3 2>&1 | tail -1 | cut -c1-200
