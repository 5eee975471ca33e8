python tools/bench_configs.py --scale 0.3 --only 3 2>&1 | tail -1 | cut -c1-200
