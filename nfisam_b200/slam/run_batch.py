"""Graph loading and incremental batching (reference: src/slam/RunBatch.py:90-346)."""
from typing import List, Tuple

from ..factors.factors import (AmbiguousDataAssociationFactor, BinaryFactor, Factor, PriorFactor,
                               SE2RelativeGaussianLikelihoodFactor)
from .graph_io import read_factor_graph_from_file
from .variables import Variable, VariableType


def graph_file_parser(data_file: str, data_format: str = "fg", prior_cov_scale=None):
    if data_format != "fg":
        raise ValueError("only the .fg format is on the path (g2o/TORO readers are out of scope): " + str(data_format))
    return read_factor_graph_from_file(data_file)


def group_nodes_factors_incrementally(nodes: List[Variable], factors: List[Factor], incremental_step: int = None,
                                      multirobot=True) -> List[Tuple[List[Variable], List[Factor]]]:
    """Time-stepped batches.  A pose is named <robot letter><step> ('X12', 'A3'): at time step t every
    robot contributes its pose t together with the factors attached to it (prior, odometry from the
    previous pose, landmark / pose observations made FROM it) and any landmark first seen there (plus
    that landmark's prior).  A batch closes every `incremental_step` time steps
    (RunBatch.py:226-336, the reference's default multirobot grouping)."""
    robots = {}
    for idx, v in enumerate(nodes):
        if v.type == VariableType.Pose:
            robots.setdefault(str(v.name)[0], {})[int(str(v.name)[1:])] = v
    last_step = max(max(steps) for steps in robots.values())
    attached = {}

    def attach(var, kind, fi):
        attached.setdefault(var, {}).setdefault(kind, []).append(fi)

    for fi, f in enumerate(factors):
        if isinstance(f, PriorFactor):
            attach(f.vars[0], "prior", fi)
        elif isinstance(f, AmbiguousDataAssociationFactor):
            attach(f.root_var, "pose_obsv" if f.child_vars[0].type == VariableType.Pose else "lmk_obsv", fi)
        elif isinstance(f, BinaryFactor):
            a, b = f.vars[0], f.vars[1]
            if a.type == b.type == VariableType.Pose:
                consecutive = str(a.name)[0] == str(b.name)[0] and int(str(b.name)[1:]) - int(str(a.name)[1:]) == 1
                if isinstance(f, SE2RelativeGaussianLikelihoodFactor) and consecutive:
                    attach(b, "odom", fi)
                else:
                    attach(a, "pose_obsv", fi)
            elif a.type == VariableType.Pose and b.type == VariableType.Landmark:
                attach(a, "lmk_obsv", fi)
            else:
                raise ValueError("Unknown factors: " + str(f))
        else:
            raise ValueError("Unknown factors: " + str(f))
    if incremental_step is None or incremental_step <= 0 or incremental_step > last_step + 1:
        incremental_step = last_step + 1
    batches, new_vars, new_factors, seen_lmk = [], [], [], set()
    for t in range(last_step + 1):
        for rid, steps in robots.items():
            if t not in steps:
                continue
            pose = steps[t]
            new_vars.append(pose)
            for kind_list in attached.get(pose, {}).values():
                new_factors += kind_list
            for fi in attached.get(pose, {}).get("lmk_obsv", []):
                for lm in factors[fi].vars[1:]:
                    if lm not in seen_lmk:
                        seen_lmk.add(lm)
                        new_vars.append(lm)
                        new_factors += attached.get(lm, {}).get("prior", [])
        if (t + 1) % incremental_step == 0 or t == last_step:
            batches.append((list(new_vars), [factors[j] for j in new_factors]))
            new_vars, new_factors = [], []
    return batches
