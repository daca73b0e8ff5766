"""hmclab_b200 -- B200-native batched HMC engine behind HMC Lab's sampler API.

Drop-in for the path ``hmclab.Samplers.HMC().sample(...)`` -> integrator ->
``Distributions.misfit/gradient`` -> accept/reject, advancing thousands of independent
Markov chains per call in hand-written sm_100a CUDA kernels reached through a C ABI
(``include/hmcb.h``).  Host code is Python; PyTorch tensors are the ``[chain x dim]``
batch container.  There is no CPU fallback: without the compiled engine library and a
CUDA device every evaluation raises.
"""
from hmclab_b200 import Distributions, MassMatrices  # noqa: F401
from hmclab_b200.Samples import Samples, combine_samples  # noqa: F401  (as hmclab/__init__.py:11)

__all__ = ["Distributions", "MassMatrices", "Samplers", "Samples", "combine_samples"]
__version__ = "0.1.0"


def __getattr__(name):
    # Samplers pulls in torch; keep `import hmclab_b200` light.
    if name == "Samplers":
        import importlib

        return importlib.import_module("hmclab_b200.Samplers")
    raise AttributeError(name)
