"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/cumicro.h declares; parameter-block layouts agree between C and ctypes."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(built):
    abi = built._abi
    lib = abi.load()
    names = abi.declared_symbols()
    assert "cumicro_bmt2m_warm_f64" in names and len(names) >= 10
    missing = [s for s in names if not hasattr(lib, s)]
    assert not missing, f"declared in include/cumicro.h but not exported: {missing}"
    assert lib.cumicro_version() == 100


def test_struct_sizes_match_c_compiler(built):
    """sizeof() of every parameter block as seen by gcc == ctypes layout."""
    abi = built._abi
    names = sorted(abi.STRUCTS)
    prog = '#include <stdio.h>\n#include "cumicro.h"\nint main(void){\n'
    for nm in names:
        prog += f'  printf("{nm} %zu\\n", sizeof({nm}));\n'
    prog += "  return 0; }\n"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", abi.INCLUDE_DIR, src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for nm in names:
        assert int(sizes[nm]) == C.sizeof(abi.STRUCTS[nm]), nm


def test_api_misuse_is_reported_without_a_gpu(built):
    """Argument validation happens before any CUDA call: status < 0 and a message."""
    abi = built._abi
    lib = abi.load()
    CMP = built.CMP
    mp = CMP.Microphysics2MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    blk = CMP.pack_2m_warm(mp, tps)
    null = C.c_void_p(0)
    st = lib.cumicro_bmt2m_warm_f64(None, C.c_int64(4), *([null] * 8), *([null] * 4), None, None)
    assert st == -1 and b"NULL" in lib.cumicro_last_error()
    st = lib.cumicro_bmt2m_warm_f64(C.byref(blk), C.c_int64(-3), *([null] * 8), *([null] * 4), None, None)
    assert st == -2
    st = lib.cumicro_bmt2m_warm_f64(C.byref(blk), C.c_int64(4), *([null] * 8), *([null] * 4), None, None)
    assert st == -1
    blk.sb.pdf_r.limited = 7
    fake = C.c_void_p(16)
    st = lib.cumicro_bmt2m_warm_f64(C.byref(blk), C.c_int64(4), *([fake] * 7), null, *([fake] * 4), None, None)
    assert st == -3 and b"limited" in lib.cumicro_last_error()
    # n == 0 is a valid empty call
    blk.sb.pdf_r.limited = 1
    st = lib.cumicro_bmt2m_warm_f64(C.byref(blk), C.c_int64(0), *([null] * 8), *([null] * 4), None, None)
    assert st == 0


def test_device_methods_refuse_cpu_arrays(built):
    import torch
    BMT, CMP = built.BMT, built.CMP
    mp = CMP.Microphysics2MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    x = torch.ones(8, dtype=torch.float64)
    with pytest.raises(built._abi.CuMicroError, match="no CPU fallback"):
        BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, x, x, x, x, x, x, x)
