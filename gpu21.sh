for cfg in "--iters 500 --lr 0.02" "--iters 1500 --lr 0.01" "--iters 1500 --lr 0.01 --samples 4000"; do
python benchmarks/solve_bench.py --robots 1 --poses 40 --landmarks 4 $cfg 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$cfg', round(j['s_per_incr_step_mean'],4), 'pose err', round(j['pose_mean_error'],2), round(j['pose_max_error'],2), 'lmk', round(j['landmark_mean_error'],2))"
done
