"""libsbn_b200: B200-native likelihood back end for phylovi/libsbn's hot path.

The package holds the CUDA kernels + C ABI (csrc/ -> lib/libsbn_b200.so) and a
thin host-side mirror of the reference interface.  Importing the package does
not load the CUDA library; the first Engine does, and fails loudly if it has
not been built.
"""
from .engine import Engine, PhyloGradient, PhyloModelSpecification, StagedBatch, TreeBatch  # noqa: F401
from .gp_engine import GPEngine, GPOperations, estimate_branch_lengths  # noqa: F401
from .site_pattern import SitePattern  # noqa: F401
from . import alignment, _capi  # noqa: F401
