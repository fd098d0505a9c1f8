"""B200-native cross-variance engine behind genomicMateSelectR's predCrossVars API."""
from ._capi import GmsError  # noqa: F401
from .engine import CrossVarEngine, trait_pairs  # noqa: F401
from .predcrossvar import (  # noqa: F401
    LDDecay, RecombFreq, crosses2predict, genmap2recombfreq, predCrossMeans, predCrossVars,
    predictCrosses, predictCrossVars,
)

__version__ = "0.1.0"
