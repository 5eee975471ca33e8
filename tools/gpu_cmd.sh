python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py --scale 0.5 --only 1 5 2>&1 | tail -2 | cut -c1-230
