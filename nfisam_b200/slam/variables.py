"""Graph variables (reference: src/slam/Variables.py).  A variable is identified, hashed and ordered
by its name; `dim` is the size of its vectorised value, rotational dims are the periodic ones."""
from enum import Enum
from typing import Hashable, List


class VariableType(Enum):
    Pose = "Pose"
    Landmark = "Landmark"
    Measurement = "Measurement"


class Variable:
    def __init__(self, name: Hashable, dim: int, variable_type: VariableType = VariableType.Pose, rotational_dims=None):
        if dim <= 0:
            raise ValueError("Dimensionality must be positive")
        rot = set(rotational_dims) if rotational_dims else set()
        if rot and not (0 <= min(rot) and max(rot) < dim):
            raise ValueError("rotational_dims is incorrect")
        self._name, self._dim, self._type, self._rot = name, int(dim), variable_type, rot
        self._hash = hash(name)          # the name never changes: hashed once (set / dict operations dominate the graph updates)

    @classmethod
    def construct_from_text(cls, line: str) -> "Variable":
        # "Variable <Pose|Landmark> <SE2|R2> <name> <truth...>"
        tok = line.strip().split()
        return _SPACES[tok[2]](name=tok[3], variable_type=VariableType(tok[1]))

    name = property(lambda self: self._name)
    dim = property(lambda self: self._dim)
    type = property(lambda self: self._type)
    rotational_dim = property(lambda self: len(self._rot))
    translational_dim = property(lambda self: self._dim - len(self._rot))

    @property
    def circular_dim_list(self) -> List[bool]:
        return [i in self._rot for i in range(self._dim)]

    @property
    def t_dim_indices(self) -> List[int]:
        return list(range(self.translational_dim))

    @property
    def R_dim_indices(self) -> List[int]:
        return list(range(self.translational_dim, self._dim))

    def __str__(self):
        return " ".join(["Variable", self._type.value, type(self).__name__.replace("Variable", ""), str(self._name)])

    __repr__ = __str__

    def __hash__(self):
        return self._hash

    def __eq__(self, other):
        return self is other or (isinstance(other, Variable) and self._name == other._name)

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        return self._name < other._name

    def __le__(self, other):
        return self._name <= other._name

    def __gt__(self, other):
        return self._name > other._name

    def __ge__(self, other):
        return self._name >= other._name


class R2Variable(Variable):
    def __init__(self, name, variable_type=VariableType.Pose):
        super().__init__(name, 2, variable_type)


class R1Variable(Variable):
    def __init__(self, name, variable_type=VariableType.Pose):
        super().__init__(name, 1, variable_type)


class SE2Variable(Variable):
    def __init__(self, name, variable_type=VariableType.Pose):
        super().__init__(name, 3, variable_type, rotational_dims={2})


_SPACES = {"R2": R2Variable, "R1": R1Variable, "SE2": SE2Variable}
