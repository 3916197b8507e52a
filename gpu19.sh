nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra 2>&1 | tail -1 > gpurun_out/bench_r1_g2.json; cut -c1-700 gpurun_out/bench_r1_g2.json
python benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 2>gpurun_out/s1.err | tail -1 > gpurun_out/solve_mr8x16_g1.json; cat gpurun_out/solve_mr8x16_g1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 2>gpurun_out/s2.err | tail -1 > gpurun_out/solve_mr8x16_g2.json; cat gpurun_out/solve_mr8x16_g2.json; tail -3 gpurun_out/s2.err
