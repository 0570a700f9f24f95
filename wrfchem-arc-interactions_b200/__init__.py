"""B200-native RRTMG SW/LW radiation hot path of the WRF-Chem ARC-interactions patch set.

Sub-modules:
  radiation  host-side mirror of RRTMG_SWRAD / RRTMG_LWRAD / rrtmg_*init (binds csrc/libarcrad.so)
  partition  j-slab column partition over GPUs + combination of per-rank domain statistics
  abi        ctypes mirror of include/arc_rad.h
  ktables    synthetic RRTMG_SW_DATA / RRTMG_LW_DATA writer (real record layout)
  synth      seeded synthetic WRF-layout columns
  decomposition  direct / semi-direct / indirect radiative effects from the domain statistics of four scenarios
"""
from . import abi, decomposition, ktables, partition, synth  # noqa: F401
from . import radiation  # noqa: F401

__all__ = ["abi", "decomposition", "ktables", "partition", "synth", "radiation"]
