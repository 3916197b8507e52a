"""Training micro driver behind profiles/r2_train_kernel.md: `python benchmarks/train_micro.py n d iters concurrent` fits `concurrent`
flows (K = 9, hidden 8) on an n x d synthetic set for `iters` Adam iterations (no early stop) and prints wall time per iteration."""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from nfisam_b200.flows import NSF_AR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 11
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
conc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rng = np.random.default_rng(7)
x = rng.standard_normal((n, d)).astype(np.float32)
for i in range(1, d):
    x[:, i] = 0.6 * x[:, i] + 0.5 * np.tanh(x[:, i - 1]) ** 2
x = (x - x.mean(0)) / x.std(0)
xd = torch.from_numpy(x).cuda()
torch.manual_seed(0)
flows = [NSF_AR(dim=d, K=9, hidden_dim=8) for _ in range(conc)]
theta0 = flows[0].flat_parameters()
streams = [torch.cuda.Stream() for _ in range(conc)]
for rep in range(3):
    for f in flows:
        f.load_flat_parameters(theta0); f.handle()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f, st in zip(flows, streams):
        f.fit_launch(xd, iters, 0.025, average_window=0 if conc == 1 else 50, loss_delta_tol=-1.0, stream=st, concurrency=conc)
    res = [f.fit_finish(pull=False) for f in flows]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"n={n} d={d} iters={iters} concurrent={conc}: {1e3*dt:.2f} ms total, {1e6*dt/iters:.2f} us/iter, loss {res[0][0][0]:.4f} -> {res[0][0][res[0][1]-1]:.4f}")
import hashlib
print("theta sha", hashlib.sha1(flows[0].flat_parameters().tobytes()).hexdigest()[:12], "hist sha", hashlib.sha1(res[0][0].tobytes()).hexdigest()[:12])
