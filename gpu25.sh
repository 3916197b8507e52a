python -m pytest tests/test_solver_gpu.py -x -q 2>&1 | tail -3
python benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 2>gpurun_out/s1.err | tail -1 > gpurun_out/solve_mr8x16_g1.json; python -c "
import json; j=json.load(open('gpurun_out/solve_mr8x16_g1.json')); print(1, j['s_per_incr_step_mean'], j['split_mean_graph_sim_train_posterior'], j['pose_mean_error'], j['landmark_mean_error'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 2>gpurun_out/s2.err | tail -1 > gpurun_out/solve_mr8x16_g2.json; python -c "
import json; j=json.load(open('gpurun_out/solve_mr8x16_g2.json')); print(2, j['s_per_incr_step_mean'], j['split_mean_graph_sim_train_posterior'], j['pose_mean_error'], j['landmark_mean_error'])"; tail -2 gpurun_out/s2.err
