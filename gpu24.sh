python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python benchmarks/solve_bench.py --robots 1 --poses 100 --landmarks 4 --iters 1500 --lr 0.01 2>gpurun_out/s1.err | tail -1 > gpurun_out/solve_manhattan100_r1.json; python -c "
import json; j=json.load(open('gpurun_out/solve_manhattan100_r1.json')); print(j['s_per_incr_step_mean'], j['s_per_incr_step_last10_mean'], j['split_mean_graph_sim_train_posterior'], j['pose_mean_error'], j['landmark_mean_error'])"
python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['incr_step']['solve_small_case1']['s_per_incr_step'])"
