from .statistics import MMDb, MMDu2, marginal_mean_cov, mmd, sample_mean  # noqa: F401
