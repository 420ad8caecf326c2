"""ncrystal_b200 -- B200-native (sm_100a) implementation of NCrystal's batched cross-section
evaluation and scatter sampling, behind NCrystal's own Scatter / C-API `*_many` interface.

The package is a thin host layer over libncrystal_b200.so (hand-written CUDA); see DESIGN.md.
"""
from .core import (Scatter, Process, Absorption, createScatter, createAbsorption, generateSource, tallyHist, kernelLaunchCount,  # noqa: F401
                   NCException, NCBadInput, NCCalcError, NCLogicError, NCFileNotFound)

__version__ = "0.1.0"
