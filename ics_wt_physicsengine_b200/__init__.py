"""B200-native batched plant-stepping engine for the wt_simulator.core hot path.

Public surface (mirrors the reference's names, see DESIGN.md / INTEGRATION.md):
  ReactorConfiguration, BoundaryConditions, ReactorState, IntegratedCSTR   (core/reactor.py)
  PlantEnsemble, EnsembleState                                             (batched form)
  calculate_pH_batch, BufferSystem, AqueousChemistry                       (core/chemistry.py)
Importing the package does not need a GPU; constructing an engine object does.
"""
from .reactor import (BoundaryConditions, EnsembleState, IntegratedCSTR, PlantEnsemble,  # noqa: F401
                      ReactorConfiguration, ReactorState)
from .chemistry import AqueousChemistry, BufferSystem, calculate_pH_batch  # noqa: F401
from . import ensembles  # noqa: F401

__version__ = "0.1.0"
