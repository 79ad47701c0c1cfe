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
    # peer-memory window: rank / size and NULL checks come before any CUDA call
    win = C.c_void_p(0)
    assert lib.cumicro_p2p_window_create(C.c_int(2), C.c_int(2), C.byref(win)) < 0 and not win.value
    assert lib.cumicro_p2p_window_create(C.c_int(0), C.c_int(17), C.byref(win)) < 0
    assert lib.cumicro_p2p_window_create(C.c_int(0), C.c_int(1), None) == -1
    assert lib.cumicro_p2p_allreduce_f64(None, fake, C.c_int(4), None) == -1 and b"window" in lib.cumicro_last_error()
    assert lib.cumicro_p2p_allreduce_f64(None, fake, C.c_int(17), None) < 0
    assert lib.cumicro_p2p_window_destroy(None) == 0


def test_device_methods_refuse_cpu_arrays(built):
    import torch
    BMT, CMP = built.BMT, built.CMP
    mp = CMP.Microphysics2MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    x = torch.ones(8, dtype=torch.float64)
    with pytest.raises(built._abi.CuMicroError, match="no CPU fallback"):
        BMT.bulk_microphysics_tendencies(BMT.Microphysics2Moment(), mp, tps, x, x, x, x, x, x, x)


def _build_c_harness(built, tmp):
    abi = built._abi
    exe = os.path.join(tmp, "c_harness")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "c_harness.c")
    subprocess.run(["gcc", "-O1", "-I", abi.INCLUDE_DIR, src, "-ldl", "-o", exe], check=True)
    return exe


def test_field_offsets_match_c_compiler(built):
    """offsetof() of the fields a foreign-language packer has to hit, as gcc lays them out, == ctypes (and with it the Julia
    mirror structs, which tests/test_julia_ext.py ties to the same header field by field)."""
    abi = built._abi
    with tempfile.TemporaryDirectory() as d:
        out = subprocess.run([_build_c_harness(built, d), "offsets"], check=True, capture_output=True, text=True).stdout
    n = 0
    for line in out.strip().splitlines():
        name, val = line.rsplit(" ", 1)
        if name.startswith("sizeof "):
            assert int(val) == C.sizeof(abi.STRUCTS[name.split()[1]])
            continue
        st, field = name.split(".")
        assert int(val) == getattr(abi.STRUCTS[st], field).offset, name
        n += 1
    assert n >= 30


def test_every_struct_every_field_offset_against_gcc(built):
    """The same for EVERY field of EVERY parameter struct (generated C program)."""
    abi = built._abi
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "cumicro.h"\nint main(void){\n'
    want = {}
    for nm, cls in sorted(abi.STRUCTS.items()):
        for fname, _ in cls._fields_:
            prog += f'  printf("{nm}.{fname} %zu\\n", offsetof({nm}, {fname}));\n'
            want[f"{nm}.{fname}"] = getattr(cls, fname).offset
    prog += "  return 0; }\n"
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "o.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "o")
        subprocess.run(["gcc", "-I", abi.INCLUDE_DIR, src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(line.rsplit(" ", 1) for line in out.strip().splitlines())
    assert len(got) == len(want) > 400
    for k, v in want.items():
        assert int(got[k]) == v, k


def test_hand_filled_c_struct_equals_the_packed_default_block(built):
    """tests/native/c_harness.c fills cumicro_params_2m_warm_f64 by hand with the reference's defaults; byte for byte it is the
    block the Python packer (and the Julia packer, same field order) produces."""
    CMP = built.CMP
    blk = CMP.pack_2m_warm(CMP.Microphysics2MParams(np.float64), CMP.ThermodynamicsParameters(np.float64))
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "c_harness.c")).read()
    prog = src.replace("int main(int argc, char** argv) {", "int main(int argc, char** argv) { if (argc > 1 && strcmp(argv[1], \"dump\") == 0) { cumicro_params_2m_warm_f64 q; fill_defaults(&q); fwrite(&q, sizeof(q), 1, stdout); return 0; }")
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "h.c")
        open(p, "w").write(prog)
        exe = os.path.join(d, "h")
        subprocess.run(["gcc", "-I", built._abi.INCLUDE_DIR, p, "-ldl", "-o", exe], check=True)
        raw = subprocess.run([exe, "dump"], check=True, capture_output=True).stdout
    mine = type(blk).from_buffer_copy(raw)
    a, b = mine.to_dict(), blk.to_dict()

    def walk(x, y, path=""):
        for k in x:
            if isinstance(x[k], dict):
                walk(x[k], y[k], path + k + ".")
            else:
                assert x[k] == y[k] or abs(x[k] - y[k]) <= 1e-15 * abs(y[k]), (path + k, x[k], y[k])
    walk(a, b)
