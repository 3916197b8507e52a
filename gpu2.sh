set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap --format=csv 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -2 > gpurun_out/bench_r1_a.json; cat gpurun_out/bench_r1_a.json | cut -c1-3000
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_a.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:nf_forward_kernel -s 1 -c 1 -o gpurun_out/prof_fwd_r1_a python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/ncu_fwd.log 2>&1
tail -2 gpurun_out/ncu_fwd.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:nf_train_kernel -s 2 -c 1 -o gpurun_out/prof_train_r1_a python bench.py --steps 1 --warmup 1 --no-cpu --samples 100000 > gpurun_out/ncu_train.log 2>&1
tail -2 gpurun_out/ncu_train.log | cut -c1-300
ls -la gpurun_out
