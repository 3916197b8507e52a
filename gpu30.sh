cat > /tmp/fk.py <<'PY'
import sys, ctypes, torch, numpy as np
sys.path.insert(0,'.')
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
for _ in range(3):
    _lib.check(lib.nfisam_factor_logpdf(arr,nd,x.data_ptr(),n,D,out.data_ptr(),None,0,st))
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:nf_factor_logpdf -s 1 -c 1 -o gpurun_out/prof_factor_r1 python /tmp/fk.py > gpurun_out/ncu_factor.log 2>&1; tail -1 gpurun_out/ncu_factor.log | cut -c1-200
cat > /tmp/tk.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
from nfisam_b200.flows import NSF_AR
torch.manual_seed(0)
f=NSF_AR(dim=12,K=9,hidden_dim=8)
x=torch.randn(1_000_000,12,device='cuda')
f.fit(x,3,0.01,average_window=0,pull=False)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:nf_train_kernel -s 1 -c 1 -o gpurun_out/prof_train_plain_r1 python /tmp/tk.py > gpurun_out/ncu_train_plain.log 2>&1; tail -1 gpurun_out/ncu_train_plain.log | cut -c1-200
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_final.json 2>/dev/null; cut -c1-400 gpurun_out/bench_r1_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1_b.csv python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1; wc -l gpurun_out/launches_r1_b.csv
python benchmarks/micro_bench.py > gpurun_out/micro_r1.jsonl 2>/dev/null; wc -l gpurun_out/micro_r1.jsonl
