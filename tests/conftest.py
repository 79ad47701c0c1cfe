import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build libcumicro.so + the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    import cumicro
    return cumicro


@pytest.fixture(scope="session")
def orc(built):
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def cuda(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_sessionfinish(session, exitstatus):
    """Float32 methods: ULP distance of the GPU results to the reference's own Float32 arithmetic (orc<float>) and of both to the
    true value, collected by cumicro.testing.assert_f32_method — written where a GPU run leaves its artefacts."""
    try:
        import json
        from cumicro import testing
        if testing.F32_REPORT:
            out = os.path.join(ROOT, "gpurun_out")
            os.makedirs(out, exist_ok=True)
            json.dump(testing.F32_REPORT, open(os.path.join(out, "f32_ulp_report.json"), "w"), indent=1)
    except Exception:
        pass
