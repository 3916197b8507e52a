nvidia-smi -L | wc -l
for N in 1 2 4 8; do
if [ $N -eq 1 ]; then
python benchmarks/solve_bench.py --robots 8 --poses 64 --landmarks 4 2>gpurun_out/s_$N.err | tail -1 > gpurun_out/solve_mr8x64_g$N.json
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N benchmarks/solve_bench.py --robots 8 --poses 64 --landmarks 4 2>gpurun_out/s_$N.err | tail -1 > gpurun_out/solve_mr8x64_g$N.json
fi
python -c "
import json; j=json.load(open('gpurun_out/solve_mr8x64_g$N.json')); print($N, round(j['s_per_incr_step_mean'],4), [round(x,4) for x in j['split_mean_graph_sim_train_posterior']], j['pose_mean_error'], j['max_level_width'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra 2>/dev/null | tail -1 > gpurun_out/bench_r1_g8.json; python -c "
import json; j=json.load(open('gpurun_out/bench_r1_g8.json')); print('bench8', j['value'], j['ms_per_step'], j['e2e']['value'])"
