python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for l in 2 3 4; do GL_HEX20_LPN=$l python tools/bench_configs.py --scale 0.2 --only 3 2>&1 | tail -1 | cut -c1-330; done
python tools/bench_configs.py --scale 0.2 --only 5 2>&1 | tail -1
