"""cProfile of a whole multi-robot solve (host side), e.g. `python benchmarks/solve_profile.py 16 50000`: where the wall time of the
steps goes outside the GPU phases."""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.solve_bench import run_solve

poses = int(sys.argv[1]) if len(sys.argv) > 1 else 16
samples = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
run_solve(robots=8, poses=4, ada_prob=0.4, iters=500, samples=2000, device=0)
run_solve(robots=8, poses=poses, ada_prob=0.4, iters=500, samples=samples, device=0)          # warm
pr = cProfile.Profile()
pr.enable()
r = run_solve(robots=8, poses=poses, ada_prob=0.4, iters=500, samples=samples, device=0, detail=True)
pr.disable()
import numpy as np
ps, sp = np.array(r["per_step"]), np.array(r["splits"])
print("mean %.4f median %.4f; split mean %s; unaccounted mean %.4f" % (ps.mean(), np.median(ps), np.round(sp.mean(0), 4), (ps - sp.sum(1)).mean()))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print("\n".join(l[:170] for l in s.getvalue().splitlines()[4:32]))
