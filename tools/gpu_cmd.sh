python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/bench_configs.py --scale 0.2 --only 4 5 2>&1 | tail -2
