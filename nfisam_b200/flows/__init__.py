from .flows import FCNN, NSF_AR, posterior_pass  # noqa: F401
from .models import NormalizingFlowModel  # noqa: F401
from .prior_dist import CustomMultivariateNormal  # noqa: F401
