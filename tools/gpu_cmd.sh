python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/bench_configs.py --scale 0.5 2>&1 | tail -4 | cut -c1-420
