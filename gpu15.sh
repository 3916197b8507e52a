python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/micro_bench.py --quick 2>/dev/null | grep '"factor"' | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(f\"{r['case']:12s} D={r['D']:2d} n={r['n']:>9d} {r['ms']:9.3f} ms {r['evals_per_s']/1e6:9.1f} M/s {r['hbm_gbs']:8.1f} GB/s\")
"
python - <<'PY'
import sys, json, ctypes, torch, numpy as np
sys.path.insert(0,'.')
from benchmarks.micro_bench import timed
from nfisam_b200 import _lib
from nfisam_b200.factors import JointFactor, _gpu
from nfisam_b200.slam.graph_io import read_factor_graph_from_file
lib=_lib.load()
nodes,truth,fs=read_factor_graph_from_file('tests/data/small_case1.fg')
j=JointFactor(fs,nodes); arr,nd=_gpu.pack_descs(j.groups())
n=10_000_000; D=22
x=torch.randn(n,D,dtype=torch.float64,device='cuda')*0.3+torch.tensor(np.concatenate([truth[v] for v in nodes]),device='cuda')
out=torch.empty(n,dtype=torch.float64,device='cuda')
st=ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
t=timed(lambda: _lib.check(lib.nfisam_factor_logpdf(arr,nd,x.data_ptr(),n,D,out.data_ptr(),None,0,st)))
print('joint14 D22 1e7:', t*1e3,'ms', 8*(D+1)*n/t*1e-9,'GB/s')
PY
python benchmarks/solve_bench.py --robots 1 --poses 100 --landmarks 4 > gpurun_out/solve_manhattan100_r1.json 2> gpurun_out/solve1.err; tail -3 gpurun_out/solve1.err; cat gpurun_out/solve_manhattan100_r1.json
python benchmarks/solve_bench.py --robots 8 --poses 16 --landmarks 4 > gpurun_out/solve_mr8x16_g1_r1.json 2> gpurun_out/solve2.err; tail -3 gpurun_out/solve2.err; cat gpurun_out/solve_mr8x16_g1_r1.json
