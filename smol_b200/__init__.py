"""smol_b200: B200-native lattice Monte-Carlo engine behind the smol.moca API.

The hot path -- per-flip cluster-expansion / Ewald delta evaluation and the Metropolis /
Wang-Landau accept-reject loop -- runs in hand-written sm_100a CUDA kernels (``csrc/``) behind
the C ABI of ``include/lmc.h``; this package is the Python host side mirroring
``smol.moca``'s Processor / Ensemble / Sampler interface.
"""
from .ensemble import Ensemble
from .processor import (ClusterDecompositionProcessor, ClusterExpansionProcessor, ClusterInteractionDistanceProcessor,
                        CompositeProcessor, CorrelationDistanceProcessor, EwaldProcessor)
from .multicell import MulticellSampler
from .sampler import Sampler
from .sublattice import Sublattice

__all__ = ["Ensemble", "Sampler", "MulticellSampler", "Sublattice", "ClusterExpansionProcessor",
           "ClusterDecompositionProcessor", "EwaldProcessor", "CompositeProcessor",
           "CorrelationDistanceProcessor", "ClusterInteractionDistanceProcessor"]
