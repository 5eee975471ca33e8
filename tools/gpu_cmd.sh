python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hex20 or tet10" 2>&1 | tail -2
python tools/bench_configs.py --scale 0.3 --only 3 2>&1 | tail -1 | cut -c1-200
GL_TILE_CPC=1 python tools/bench_configs.py --scale 0.3 --only 3 2>&1 | tail -1 | cut -c1-200
