"""Per-step cProfile of the 8-robot x 64-pose solve (profiles/r2_scheduler.md): prints the host-side profile of the slow steps.
This is how the per-launch cudaFuncSetAttribute stall was found."""
import cProfile, io, pstats, sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
from nfisam_b200.slam.run_batch import group_nodes_factors_incrementally
from nfisam_b200.slam.synthetic import make_manhattan_range_graph
from benchmarks.solve_bench import run_solve
run_solve(robots=8, poses=4, ada_prob=0.4, iters=500, samples=2000, device=0)
nodes, truth, factors = make_manhattan_range_graph(robots=8, poses=64, landmarks=4, ada_prob=0.4, seed=0)
steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=1)
solver = NFiSAM(NFiSAMArgs(num_knots=9, flow_iterations=500, local_sample_num=2000, learning_rate=0.02, hidden_dim=8,
                           posterior_sample_num=1000, elimination_method="pose_first", deterministic_cliques=True, seed=0, device=0))
shown = 0
for k, (sn, sf) in enumerate(steps):
    for v in sn: solver.add_node(v)
    for f in sf: solver.add_factor(f)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    solver.update_physical_and_working_graphs()
    solver.incremental_inference()
    torch.cuda.synchronize()
    pr.disable()
    dt = time.perf_counter() - t0
    if dt > 0.06 and k > 5 and shown < 3:
        shown += 1
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(12)
        print('==== step', k, 'took %.4f' % dt)
        print('\n'.join(l[:160] for l in s.getvalue().splitlines()[4:26]))
print('done')
