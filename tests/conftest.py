import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import bindings
    return bindings.oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference compiled in place (oracle/_ref).  Absent on a box without /root/reference and without the built .so."""
    from oracle import bindings
    r = bindings.ref()
    if r is None:
        pytest.skip("oracle/_ref/libphmm_ref.so not built (needs /root/reference)")
    return r


@pytest.fixture(scope="session")
def engine():
    """The engine the parity tests run on: every batch, however small, goes through the DEVICE epilogue (the default, AUTO,
    keeps batches under 32 768 pairs on the host; test_gpu_parity.py has a test of its own for that)."""
    from rovaca_b200 import PairHmmEngine
    from rovaca_b200.pairhmm import EPILOGUE_DEVICE
    eng = PairHmmEngine(epilogue=EPILOGUE_DEVICE)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_nofma():
    from rovaca_b200 import PairHmmEngine, FMA_NONE
    eng = PairHmmEngine(fma_mode=FMA_NONE)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_host_epilogue():
    """An engine that keeps the epilogue (log10, normalize, filter, transpose, best haplotype) on the host."""
    from rovaca_b200 import PairHmmEngine
    from rovaca_b200.pairhmm import EPILOGUE_HOST
    eng = PairHmmEngine(epilogue=EPILOGUE_HOST)
    yield eng
    eng.close()
