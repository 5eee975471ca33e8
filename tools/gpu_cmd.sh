python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pat in 2 1 0; do
GL_HEX8_PATTERN=$pat python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_h8v4_p$pat.json 2> gpurun_out/bench_h8v4.err; tail -3 gpurun_out/bench_h8v4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h8v4_p$pat.json')); print($pat, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['checks'])"
done
