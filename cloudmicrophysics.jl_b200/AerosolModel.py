"""Mirror of ``CloudMicrophysics.AerosolModel`` (src/AerosolModel.jl:26-103): aerosol size-
distribution modes and the distribution container (host-side parameter objects)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class Mode_B:
    """Mode_B (AerosolModel.jl:26-45): B-parameter (Abdul-Razzak & Ghan 2000) description."""
    r_dry: float
    stdev: float
    N: float
    mass_mix_ratio: Tuple[float, ...]
    soluble_mass_frac: Tuple[float, ...]
    osmotic_coeff: Tuple[float, ...]
    molar_mass: Tuple[float, ...]
    dissoc: Tuple[float, ...]
    aerosol_density: Tuple[float, ...]


@dataclass(frozen=True)
class Mode_κ:
    """Mode_κ (AerosolModel.jl:61-76): kappa (Petters & Kreidenweis 2007) description."""
    r_dry: float
    stdev: float
    N: float
    vol_mix_ratio: Tuple[float, ...]
    mass_mix_ratio: Tuple[float, ...]
    molar_mass: Tuple[float, ...]
    kappa: Tuple[float, ...]


Mode_kappa = Mode_κ


@dataclass(frozen=True)
class AerosolDistribution:
    """AerosolDistribution (AerosolModel.jl:93-103): a tuple of modes of one kind."""
    modes: tuple

    def __post_init__(self):
        if len({type(m) for m in self.modes}) > 1:
            raise TypeError("all modes of an AerosolDistribution must be of the same type")


def n_modes(ad) -> int:
    return len(ad.modes)


def n_components(mode) -> int:
    return len(mode.mass_mix_ratio)
