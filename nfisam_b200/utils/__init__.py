from .statistics import MMDb, MMDu2, mmd  # noqa: F401
