python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 2>gpurun_out/s1.err | tail -1 > gpurun_out/solve_mr8x16_g1.json; python -c "
import json; j=json.load(open('gpurun_out/solve_mr8x16_g1.json')); print(j['s_per_incr_step_mean'], j['split_mean_graph_sim_train_posterior'], j['pose_mean_error'])"
python benchmarks/solve_bench.py --robots 1 --poses 100 --landmarks 4 2>gpurun_out/s1.err | tail -1 > gpurun_out/solve_manhattan100_r1.json; python -c "
import json; j=json.load(open('gpurun_out/solve_manhattan100_r1.json')); print(j['s_per_incr_step_mean'], j['split_mean_graph_sim_train_posterior'], j['pose_mean_error'], j['landmark_mean_error'])"
