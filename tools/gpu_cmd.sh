python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/bench_configs.py --scale 1.0 2>gpurun_out/configs_full.err | tee gpurun_out/configs_full.jsonl | cut -c1-300
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -2 gpurun_out/bench_r1_final.err; cut -c1-200 gpurun_out/bench_r1_final.json
