"""`.fg` text format reader / writer (reference: src/slam/FactorGraphSimulator.py:20-74).

    Variable <Pose|Landmark> <SE2|R2> <name> <truth...>
    Factor <ClassName> ...        (dispatched to <ClassName>.construct_from_text)
"""
from typing import Dict, List, Tuple

import numpy as np

from ..factors.factors import Factor
from .variables import Variable


def read_factor_graph_from_file(path: str) -> Tuple[List[Variable], Dict[Variable, np.ndarray], List[Factor]]:
    nodes, truth, factors = [], {}, []
    with open(path) as fh:
        lines = [ln.strip() for ln in fh if ln.strip()]
    for ln in lines:
        tok = ln.split()
        if tok[0] == "Variable":
            var = Variable.construct_from_text(ln)
            nodes.append(var)
            truth[var] = np.array([float(t) for t in tok[4:4 + var.dim]])
    for ln in lines:
        if ln.split()[0] == "Factor":
            factors.append(Factor.construct_from_text(ln, nodes))
    return nodes, truth, factors


def factor_graph_to_string(variables, factors, var_truth=None) -> str:
    out = []
    for v in variables:
        line = str(v)
        if var_truth is not None and v in var_truth:
            line += " " + " ".join(str(t) for t in var_truth[v])
        out.append(line)
    out += [str(f) for f in factors]
    return "\n".join(out)
