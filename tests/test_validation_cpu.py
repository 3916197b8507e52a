"""Host logic of the validation-split path (training_set_frac < 1) through the oracle backend."""
import os

import numpy as np
import torch

from tests.oracle_backend import oracle_backend

HERE = os.path.dirname(os.path.abspath(__file__))


def test_solver_with_validation_split_runs_and_stops_early():
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.slam.run_batch import graph_file_parser, group_nodes_factors_incrementally

    nodes, truth, factors = graph_file_parser(os.path.join(HERE, "data", "small_case1.fg"))
    steps = group_nodes_factors_incrementally(nodes, factors, incremental_step=3)
    np.random.seed(0)
    torch.manual_seed(0)
    with oracle_backend():
        solver = NFiSAM(NFiSAMArgs(num_knots=5, flow_iterations=60, local_sample_num=300, posterior_sample_num=100,
                                   learning_rate=0.05, training_set_frac=0.7, validation_interval=5, slower_stop_rate=1.5))
        for sn, sf in steps:
            for v in sn:
                solver.add_node(v)
            for f in sf:
                solver.add_factor(f)
            solver.update_physical_and_working_graphs()
            cur = solver.incremental_inference()
    assert len(cur) == 8
    curves = list(solver._temp_training_loss.values())
    assert curves and all(len(c) == 60 for c in curves)
