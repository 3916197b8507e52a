"""Factor graph + symbolic elimination into a Bayes tree (reference: src/slam/FactorGraph.py)."""
from typing import Dict, Iterable, List, Set

from ..factors.factors import Factor
from .bayes_tree import BayesTree, BayesTreeNode
from .variables import Variable


class FactorGraph:
    def __init__(self):
        self._vars: List[Variable] = []
        self._factors: List[Factor] = []
        self._var_set: Set[Variable] = set()
        self._incidence = None            # (variable -> insertion position, variable -> indices of its factors), built lazily
        self._adjacency = None            # built lazily: sub-graphs are created far more often than eliminated

    vars = property(lambda self: self._vars)
    factors = property(lambda self: self._factors)

    def add_node(self, var: Variable) -> "FactorGraph":
        if var in self._var_set:
            raise KeyError("The node has already existed in the graph")
        if self._incidence is not None:
            self._incidence[0][var] = len(self._vars)
            self._incidence[1][var] = []
        self._vars.append(var)
        self._var_set.add(var)
        self._adjacency = None
        return self

    def add_factor(self, factor: Factor) -> "FactorGraph":
        for v in factor.vars:
            if v not in self._var_set:
                raise KeyError(f"factor touches a variable that is not in the graph: {v.name}")
        if self._incidence is not None:
            for v in factor.vars:
                self._incidence[1][v].append(len(self._factors))
        self._factors.append(factor)
        self._adjacency = None
        return self

    def _incidence_maps(self):
        """(variable -> insertion position, variable -> indices of the factors that touch it); kept up to date by add_node /
        add_factor once built (the physical graph only ever grows), rebuilt after an in-place removal."""
        if self._incidence is None:
            index = {v: k for k, v in enumerate(self._vars)}
            incident = {v: [] for v in self._vars}
            for fi, f in enumerate(self._factors):
                for v in f.vars:
                    incident[v].append(fi)
            self._incidence = (index, incident)
        return self._incidence

    def take_clique(self, clique: BayesTreeNode) -> List[Factor]:
        """Symbolic elimination of a clique IN PLACE: removes the clique's frontal variables and every factor that lies inside
        the clique, and returns those factors in insertion order (= get_clique_factor_graph(clique).factors followed by
        eliminate_clique_variables(clique, None), FactorGraph.py:230-259, without building two new graphs per clique)."""
        cv = clique.vars
        frontal = clique.frontal
        taken, kept = [], []
        for f in self._factors:
            (taken if cv.issuperset(f.vars) else kept).append(f)
        self._factors = kept
        self._vars = [v for v in self._vars if v not in frontal]
        self._var_set -= frontal
        self._adjacency = None
        self._incidence = None
        return taken

    @property
    def _neighbors(self) -> Dict[Variable, Set[Variable]]:
        if self._adjacency is None:
            adj = {v: set() for v in self._vars}
            for f in self._factors:
                vs = list(f.vars)
                for a in vs:
                    for b in vs:
                        if a != b:
                            adj[a].add(b)
            self._adjacency = adj
        return self._adjacency

    def get_neighbors_in_factor_graph(self, key: Variable) -> Set[Variable]:
        return self._neighbors[key]

    def get_bayes_tree(self, ordering: List[Variable]) -> BayesTree:
        """Symbolic elimination along `ordering`: the separator of a variable is its neighbour set at
        elimination time, which then becomes a clique (FactorGraph.py:70-92, 172-202)."""
        adj = {v: set(n) for v, n in self._neighbors.items()}
        parents = {}
        for v in ordering:
            sep = set(adj[v])
            for u in sep:
                adj[u].discard(v)
                adj[u] |= sep - {u}
            adj[v] = set()
            parents[v] = sep
        tree = BayesTree(frontal=ordering[-1])
        tree.reverse_elimination_order = ordering[::-1]
        for v in ordering[:-1][::-1]:
            tree.add_node(frontal=v, parents=parents[v])
        return tree

    def _subgraph(self, keep_var, keep_factor, extra: Iterable[Factor] = ()) -> "FactorGraph":
        g = FactorGraph()
        for v in self._vars:
            if keep_var(v):
                g.add_node(v)
        for f in self._factors:
            if keep_factor(f):
                g.add_factor(f)
        for f in extra:
            if f is not None:
                g.add_factor(f)
        return g

    def get_sub_factor_graph_with_prior(self, variables: Set[Variable], sub_trees: List[BayesTree],
                                        clique_prior_dict: Dict[BayesTreeNode, Factor]) -> "FactorGraph":
        """Factors among `variables` that are not already summarised by an untouched subtree, plus the
        separator factors of those subtrees (FactorGraph.py:204-228)."""
        roots = [t.root.vars for t in sub_trees]
        # only the factors incident to the affected variables are looked at (the physical graph grows with the trajectory,
        # the affected part does not); variables and factors keep their insertion order, as a scan of the whole graph would
        var_index, incident = self._incidence_maps()
        candidates = set()
        for v in variables:
            candidates.update(incident[v])
        g = FactorGraph()
        for v in sorted(variables, key=var_index.__getitem__):
            g.add_node(v)
        for fi in sorted(candidates):
            f = self._factors[fi]
            fv = set(f.vars)
            if fv.issubset(variables) and not any(fv.issubset(r) for r in roots):
                g.add_factor(f)
        for t in sub_trees:
            prior = clique_prior_dict.get(t.root)
            if prior is not None:
                g.add_factor(prior)
        return g

    def eliminate_clique_variables(self, clique: BayesTreeNode, new_factor: Factor) -> "FactorGraph":
        """Drop the clique's frontal variables and every factor inside the clique; add its separator
        factor (FactorGraph.py:230-247)."""
        cv = clique.vars
        return self._subgraph(lambda v: v not in clique.frontal, lambda f: not set(f.vars).issubset(cv), [new_factor])

    def get_clique_factor_graph(self, clique: BayesTreeNode) -> "FactorGraph":
        cv = clique.vars
        return self._subgraph(lambda v: v in cv, lambda f: set(f.vars).issubset(cv))
