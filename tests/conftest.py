import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def lib(cuda_device):
    from bimodalattack_b200 import native

    return native.load()


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    """On a failing GPU test, show the device-side watchdog word (csrc/common.cuh): a pipeline wait
    that timed out leaves its barrier code there even after the CUDA context has died."""
    outcome = yield
    rep = outcome.get_result()
    if rep.when == "call" and rep.failed and item.get_closest_marker("gpu") is not None:
        try:
            import ctypes
            from bimodalattack_b200 import native
            code = ctypes.c_uint(0)
            native.load().bma_watchdog_code(ctypes.byref(code))
            rep.sections.append(("bimodal_b200 watchdog", f"code 0x{code.value:x} (0 = no pipeline wait timed out)"))
        except Exception as exc:  # noqa: BLE001
            rep.sections.append(("bimodal_b200 watchdog", f"unavailable: {exc}"))
