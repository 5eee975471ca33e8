python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/bench_configs.py --scale 0.5 --only 1 5 2>&1 | tail -2 | cut -c1-330
GL_DISABLE_SMALL2D=1 python tools/bench_configs.py --scale 0.5 --only 1 5 2>&1 | tail -2 | cut -c1-200
