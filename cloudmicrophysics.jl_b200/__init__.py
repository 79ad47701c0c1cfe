"""cumicro — B200-native bulk cloud-microphysics tendencies (host-side mirror of
the CloudMicrophysics.jl module API for the hot path).  See DESIGN.md.

Module names follow the reference: ``CMP`` (parameters), ``BMT``
(BulkMicrophysicsTendencies), ``CM2`` (Microphysics2M).  The array methods need the CUDA
library (``libcumicro.so``) and a CUDA device; nothing here falls back to the CPU."""
from . import _abi  # noqa: F401
from . import parameters as CMP  # noqa: F401
from . import testing  # noqa: F401


def __getattr__(name):  # torch-dependent modules are imported on first use
    import importlib
    lazy = {"BulkMicrophysicsTendencies": "BulkMicrophysicsTendencies", "BMT": "BulkMicrophysicsTendencies",
            "Microphysics2M": "Microphysics2M", "CM2": "Microphysics2M",
            "Microphysics1M": "Microphysics1M", "CM1": "Microphysics1M",
            "MicrophysicsNonEq": "MicrophysicsNonEq", "CMNonEq": "MicrophysicsNonEq",
            "CloudDiagnostics": "CloudDiagnostics", "CMD": "CloudDiagnostics",
            "AerosolActivation": "AerosolActivation", "AA": "AerosolActivation",
            "AerosolModel": "AerosolModel", "AM": "AerosolModel",
            "IceNucleation": "IceNucleation", "IN": "IceNucleation", "fused": "fused", "collective": "collective",
            "EmulatorModels": "EmulatorModels",
            "P3Scheme": "P3Scheme", "P3": "P3Scheme", "CMP3": "parameters_p3"}
    if name in lazy:
        return importlib.import_module("." + lazy[name], __name__)
    raise AttributeError(name)
