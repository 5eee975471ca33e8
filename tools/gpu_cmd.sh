python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_r1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_r1c.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_r1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:hex8 -s 3 -c 1 -o gpurun_out/prof_h8_r1c -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_h8_r1c.log 2>&1
tail -1 gpurun_out/ncu_h8_r1c.log
