"""nfisam_b200 -- B200 (sm_100a) implementation of NF-iSAM's per-clique normalizing-flow hot path.

The compute lives in ``libnfisam_b200.so`` (hand-written CUDA, C ABI in ``include/nfisam_b200.h``);
this package is the host-side mirror of the reference's Python interfaces for that path:

    nfisam_b200.flows      NSF_AR / FCNN / NormalizingFlowModel / CustomMultivariateNormal
                           (reference: src/flows/{flows,models,prior_dist}.py)
    nfisam_b200.factors    batched factor log-likelihoods (reference: src/factors/Factors.py)
    nfisam_b200.slam       NFiSAM solver plugin + clique scheduler (reference: src/slam/NFiSAM.py)

There is no CPU fallback: importing is cheap, but every compute call raises if the CUDA library
or a CUDA device is missing.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
