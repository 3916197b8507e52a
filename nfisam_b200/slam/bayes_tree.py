"""Bayes tree bookkeeping read by the clique scheduler (reference: src/slam/BayesTree.py).

Same public surface as the reference (BayesTreeNode: frontal / separator / children / parent,
dim properties, equality by variable sets; BayesTree: add_node, clique_nodes, leaves,
get_affected_vars_and_partial_bayes_trees, clique_variable_pattern, clique_ordering,
append_child_bayes_trees, __copy__) with one deliberate difference: every traversal runs in a
deterministic order (children are kept in insertion order and a new variable is attached to the
clique whose frontal set holds its earliest-eliminated parent), whereas the reference iterates
Python sets (BayesTree.py:215-231), which makes its tree shape depend on PYTHONHASHSEED."""
from typing import Iterable, List, Set, Tuple, Union

from .variables import Variable


class BayesTreeNode:
    def __init__(self, frontal: Union[Variable, Set[Variable]], separator: Set[Variable] = None,
                 children=None, parent: "BayesTreeNode" = None):
        if isinstance(frontal, Variable):
            self.frontal = {frontal}
        elif isinstance(frontal, set):
            self.frontal = frontal
        else:
            raise ValueError("The frontal must be either the set of all frontal variables, or a frontal variable")
        self.separator = separator if separator else set()
        self.parent = parent
        self.children: List["BayesTreeNode"] = list(children) if children else []
        self._hash = None       # memoised __hash__: (len(frontal), len(separator), value); dropped by add_frontal

    # -- structure -------------------------------------------------------------------------
    def append_child(self, child: "BayesTreeNode") -> "BayesTreeNode":
        if not any(c is child for c in self.children):
            self.children.append(child)
        child.parent = self
        return self

    def create_child(self, frontal: Variable, separator: Set[Variable] = None) -> "BayesTreeNode":
        child = BayesTreeNode(frontal=frontal, separator=set(separator) if separator else set())
        self.append_child(child)
        return child

    def add_frontal(self, frontal: Variable) -> "BayesTreeNode":
        self.frontal.add(frontal)
        self._hash = None
        return self

    def remove_child(self, child: "BayesTreeNode") -> "BayesTreeNode":
        self.children = [c for c in self.children if c is not child]
        child.parent = None
        return self

    # -- queries ---------------------------------------------------------------------------
    is_leaf = property(lambda self: len(self.children) == 0)
    is_root = property(lambda self: self.parent is None)
    vars = property(lambda self: self.frontal | self.separator)
    num_vars = property(lambda self: len(self.frontal) + len(self.separator))
    dim = property(lambda self: sum(v.dim for v in self.frontal | self.separator))
    separator_dim = property(lambda self: sum(v.dim for v in self.separator))
    frontal_dim = property(lambda self: sum(v.dim for v in self.frontal))

    def copy_without_parents_children(self) -> "BayesTreeNode":
        node = BayesTreeNode(frontal=set(self.frontal), separator=set(self.separator))
        node._hash = self._hash
        return node

    def deep_copy(self) -> "BayesTreeNode":
        """Copy of the subtree rooted here (parent link of the copy is None).  Iterative: chain-shaped trees are as deep
        as they have cliques (a 1500-pose trajectory would overflow Python's recursion limit)."""
        root = self.copy_without_parents_children()
        stack = [(self, root)]
        while stack:
            src, dst = stack.pop()
            for ch in src.children:
                cp = ch.copy_without_parents_children()
                cp.parent = dst
                dst.children.append(cp)
                stack.append((ch, cp))
        return root

    def __eq__(self, other) -> bool:
        return isinstance(other, BayesTreeNode) and self.frontal == other.frontal and self.separator == other.separator

    def __hash__(self) -> int:
        # same value as the reference's (BayesTree.py:157-159); memoised because the solver looks cliques up in dicts once per
        # clique and pass (4.5 us each: 2.3 ms per posterior pass on a 500-clique tree).  The sizes guard the memo against
        # direct mutation of the public sets.
        memo = self._hash
        if memo is not None and memo[0] == len(self.frontal) and memo[1] == len(self.separator):
            return memo[2]
        value = hash((tuple(sorted(str(v.name) for v in self.separator)), tuple(sorted(str(v.name) for v in self.frontal))))
        self._hash = (len(self.frontal), len(self.separator), value)
        return value

    def __str__(self) -> str:
        names = lambda s: "{" + ", ".join(sorted(str(v.name) for v in s)) + "}"  # noqa: E731
        return f"BayesTreeNode(frontal={names(self.frontal)}, separator={names(self.separator)}, children={len(self.children)})"

    __repr__ = __str__


class BayesTree:
    def __init__(self, root_clique: BayesTreeNode = None, frontal: Variable = None):
        if root_clique is not None:
            self.root = root_clique
            for child in root_clique.children:
                child.parent = root_clique
        elif frontal is not None:
            self.root = BayesTreeNode(frontal=frontal)
        else:
            raise ValueError("Either the root clique or a root frontal variable needs to be specified")
        self.reverse_elimination_order = None

    def _walk(self) -> List[BayesTreeNode]:
        """Breadth-first list of cliques, root first (deterministic)."""
        out, queue = [], [self.root]
        while queue:
            c = queue.pop(0)
            out.append(c)
            queue.extend(c.children)
        return out

    clique_nodes = property(lambda self: set(self._walk()))
    leaves = property(lambda self: {c for c in self._walk() if c.is_leaf})
    frontal_vars = property(lambda self: set().union(*[c.frontal for c in self._walk()]))

    def clique_ordering(self) -> List[BayesTreeNode]:
        """BFS list, root first; the solver pops from the end (leaves -> root), BayesTree.py:375-384."""
        return self._walk()

    def levels(self) -> List[List[BayesTreeNode]]:
        """Cliques grouped by height above the leaves: levels()[0] are cliques without children,
        levels()[h] have all children in lower levels.  Cliques inside one level are mutually
        independent during training -- the unit of the clique-parallel schedule."""
        cliques = self._walk()                 # breadth first: children come after their parent
        height = {}
        for c in reversed(cliques):            # ... so every child is done before its parent (no recursion)
            height[id(c)] = 1 + max(height[id(ch)] for ch in c.children) if c.children else 0
        out = [[] for _ in range(height[id(self.root)] + 1)]
        for c in cliques:
            out[height[id(c)]].append(c)
        return out

    def add_node(self, frontal: Variable, parents: Set[Variable] = None) -> "BayesTree":
        """Insert a variable (in reverse elimination order) whose Bayes-net parents are `parents`."""
        parents = set(parents) if parents else set()
        cliques = self._walk()
        target = None
        if parents and self.reverse_elimination_order is not None:
            pos = {v: k for k, v in enumerate(self.reverse_elimination_order)}
            first = max(parents, key=lambda v: pos[v])          # the parent eliminated earliest
            for c in cliques:
                if first in c.frontal and parents.issubset(c.vars):
                    target = c
                    break
        if target is None:
            for c in cliques:
                if parents.issubset(c.vars):
                    target = c
                    break
        if target is None:
            raise ValueError("no clique contains the parents of " + str(frontal.name))
        if len(parents) == target.num_vars:
            target.add_frontal(frontal)
        else:
            target.create_child(frontal, parents)
        return self

    def append_clique(self, clique: BayesTreeNode, parent_clique: BayesTreeNode) -> "BayesTree":
        parent_clique.append_child(clique)
        return self

    def append_child_bayes_tree(self, child_tree: "BayesTree", candidates: List[BayesTreeNode] = None) -> "BayesTree":
        """Hang `child_tree` below the first clique (breadth-first order, or the given candidate list) that contains its
        root's separator."""
        for attach_point in (candidates if candidates is not None else self._walk()):
            if child_tree.root.separator.issubset(attach_point.vars):
                attach_point.append_child(child_tree.root)
                break
        return self

    def append_child_bayes_trees(self, child_trees: Iterable["BayesTree"], among_current: bool = False) -> "BayesTree":
        """among_current: only the cliques the tree holds NOW are attach points (the incremental update appends the
        untouched subtrees of the previous tree to the re-eliminated part: their separators live in that part, and the
        scan then does not grow with the size of the appended subtrees)."""
        candidates = self._walk() if among_current else None
        for t in child_trees:
            self.append_child_bayes_tree(t, candidates)
        return self

    def __copy__(self) -> "BayesTree":
        new_tree = BayesTree(root_clique=self.root.deep_copy())
        if self.reverse_elimination_order:
            new_tree.reverse_elimination_order = list(self.reverse_elimination_order)
        return new_tree

    def frontal_owner_map(self) -> dict:
        """variable -> the clique that holds it as a frontal variable."""
        owner = {}
        for c in self._walk():
            for v in c.frontal:
                owner[v] = c
        return owner

    def get_affected_vars_and_partial_bayes_trees(self, vars: Set[Variable]) -> Tuple[Set[Variable], List["BayesTree"]]:
        """Cliques holding one of `vars` as frontal, plus all their ancestors, are affected; every
        maximal unaffected subtree is returned as its own (copied) tree (BayesTree.py:310-356)."""
        affected_vars, sub_trees, _ = self.split_affected(vars, reuse_nodes=False)
        return affected_vars, sub_trees

    def split_affected(self, vars: Set[Variable], owner: dict = None, reuse_nodes: bool = True):
        """The incremental form of the above: (affected variables, untouched subtrees, affected cliques).  `owner` is a
        frontal_owner_map the caller maintains across steps (None: built here, one walk).  With reuse_nodes the untouched
        subtrees are DETACHED from this tree and handed over as they are, like the reference's shallow subtree copy
        (BayesTree.py:336-356): the work per step then scales with the affected part, not with the trajectory length --
        this tree must not be used afterwards."""
        if owner is None:
            owner = self.frontal_owner_map()
        affected = set()
        for v in vars:
            c = owner.get(v)
            while c is not None and id(c) not in affected:
                affected.add(id(c))
                c = c.parent
        affected_vars, sub_trees, removed = set(), [], []
        stack = [self.root]
        while stack:
            c = stack.pop()
            affected_vars |= c.frontal
            removed.append(c)
            for ch in c.children:
                if id(ch) in affected:
                    stack.append(ch)
                elif reuse_nodes:
                    ch.parent = None
                    sub_trees.append(BayesTree(root_clique=ch))
                else:
                    sub_trees.append(BayesTree(root_clique=ch.deep_copy()))
        if id(self.root) not in affected:
            # nothing touched: the reference still treats the root path as affected via the union below
            affected_vars = set(self.root.frontal)
        return affected_vars, sub_trees, removed

    def clique_variable_pattern(self, clique: BayesTreeNode) -> List[Variable]:
        """[separator variables, frontal variables], each in reverse elimination order (BayesTree.py:358-373)."""
        pos = {v: k for k, v in enumerate(self.reverse_elimination_order)}
        return sorted(clique.separator, key=lambda v: pos[v]) + sorted(clique.frontal, key=lambda v: pos[v])

    def __str__(self) -> str:
        return "BayesTree{" + ", ".join(str(c) for c in self._walk()) + "}"
