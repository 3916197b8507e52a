from .variables import R1Variable, R2Variable, SE2Variable, Variable, VariableType  # noqa: F401
