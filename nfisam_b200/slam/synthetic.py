"""Synthetic Manhattan-world range-SLAM graphs for the large BASELINE configs (100+ poses single
robot; R robots x T poses with shared landmarks for the clique-parallel runs).

Own generator with the reference simulator's ingredients (src/manhattan_world_with_range:
grid moves with 90-degree turns, SE(2) odometry noise, range-only landmark observations, optional
ambiguous data association between the true landmark and a decoy): pose names are
<robot letter><time step> ('A0', 'B17'), the naming the reference's multi-robot batching keys on
(src/slam/RunBatch.py:226-336); every robot gets a prior on its first pose (a clique without a prior
cannot be sampled ancestrally)."""
from typing import Dict, List, Tuple

import numpy as np

from ..factors.factors import (AmbiguousDataAssociationFactor, Factor, SE2R2RangeGaussianLikelihoodFactor,
                               SE2RelativeGaussianLikelihoodFactor, UnarySE2ApproximateGaussianPriorFactor)
from ..factors.geometry import SE2Pose, se2_compose, se2_exp
from .variables import R2Variable, SE2Variable, Variable, VariableType


def make_manhattan_range_graph(robots: int = 1, poses: int = 16, landmarks: int = 4, cell: float = 10.0,
                               range_sigma: float = 2.0, odom_sigmas=(0.2, 0.04, 0.02), prior_sigmas=(0.02, 0.02, 0.002),
                               max_range: float = 1e9, ada_prob: float = 0.0, seed: int = 0, ranges_per_pose: int = 2
                               ) -> Tuple[List[Variable], Dict[Variable, np.ndarray], List[Factor]]:
    """Returns (nodes, truth, factors) like read_factor_graph_from_file.  Every pose measures the range
    to its `ranges_per_pose` nearest landmarks (with probability `ada_prob` an association is ambiguous
    between the true landmark and a random other one); at time 0 each robot ranges to every landmark so
    that all landmarks are tied to a prior-connected pose."""
    rng = np.random.default_rng(seed)
    side = int(np.ceil(np.sqrt(max(robots, 1))))
    extent = cell * max(4, int(np.sqrt(poses)) + 2)
    lmk_xy = rng.uniform(-0.5 * extent, 0.5 * extent + side * extent * 0.5, size=(landmarks, 2))
    lmks = [R2Variable(f"L{k + 1}", VariableType.Landmark) for k in range(landmarks)]
    truth: Dict[Variable, np.ndarray] = {}
    nodes: List[Variable] = []
    factors: List[Factor] = []
    odom_cov = np.diag(np.square(odom_sigmas))
    prior_cov = np.diag(np.square(prior_sigmas))
    robot_nodes = []
    for r in range(robots):
        letter = chr(ord("A") + r) if robots > 1 else "X"
        start = np.array([(r % side) * extent * 0.5, (r // side) * extent * 0.5, rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2])])
        pose = start.copy()
        seq = []
        for t in range(poses):
            var = SE2Variable(f"{letter}{t}", VariableType.Pose)
            if t > 0:
                turn = rng.choice([0.0, np.pi / 2, -np.pi / 2], p=[0.6, 0.2, 0.2])
                step = np.array([cell, 0.0, turn])
                new_pose = se2_compose(pose, step)[0]
                noise = se2_exp(rng.standard_normal(3) * np.asarray(odom_sigmas))[0]
                meas = se2_compose(step, noise)[0]
                factors.append(SE2RelativeGaussianLikelihoodFactor(seq[-1], var, SE2Pose(*meas), odom_cov))
                pose = new_pose
            else:
                factors.append(UnarySE2ApproximateGaussianPriorFactor(var, SE2Pose(*pose), prior_cov))
            truth[var] = pose.copy()
            seq.append(var)
        robot_nodes.append(seq)
    seen = set()
    ordered_nodes: List[Variable] = []
    for t in range(poses):
        for seq in robot_nodes:
            var = seq[t]
            ordered_nodes.append(var)
            d = np.linalg.norm(lmk_xy - truth[var][:2], axis=1)
            targets = list(range(landmarks)) if t == 0 else [int(k) for k in np.argsort(d)[:max(1, ranges_per_pose)]]
            for k in targets:
                if d[k] > max_range:
                    continue
                obs = float(d[k] + rng.standard_normal() * range_sigma)
                if lmks[k] not in seen:
                    seen.add(lmks[k])
                    ordered_nodes.append(lmks[k])
                    truth[lmks[k]] = lmk_xy[k].copy()
                others = [j for j in range(landmarks) if j != k and lmks[j] in seen]
                if t > 0 and others and rng.random() < ada_prob:
                    j = int(rng.choice(others))
                    factors.append(AmbiguousDataAssociationFactor(var, [lmks[k], lmks[j]], np.array([0.5, 0.5]),
                                                                  SE2R2RangeGaussianLikelihoodFactor, obs, range_sigma))
                else:
                    factors.append(SE2R2RangeGaussianLikelihoodFactor(var, lmks[k], obs, range_sigma))
    nodes = ordered_nodes
    return nodes, truth, factors
