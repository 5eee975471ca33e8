python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py --scale 0.5 --only 4 2>&1 | tail -1 | cut -c1-330
