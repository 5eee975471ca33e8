free -g | head -2; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-layers 400 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8.err; tail -3 gpurun_out/bench_n8.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n8b.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['e2e']['workload'][:40], d['clocks'])"
