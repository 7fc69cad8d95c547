"""particlesim_b200 — B200 (sm_100a) implementation of ParticleSim's force hot path.

Only what the path needs: `csrc/` (CUDA kernels + the C ABI of include/psim_b200.h), the ctypes
loader and the host-side mirror of the reference's Quadtree / CellList / forces interface.
"""
from ._lib import PsimError, default_config, default_species_table, load  # noqa: F401
from .simulation import (Bodies, CellList, Quadtree, SimConfig, Simulation,  # noqa: F401
                         coulomb_constant, forces)

__all__ = ["Bodies", "CellList", "Quadtree", "SimConfig", "Simulation", "forces", "PsimError",
           "default_config", "default_species_table", "load", "coulomb_constant"]
