import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    if not os.path.exists(os.path.join(oracle_lib.ORACLE_DIR, "liboracle.so")):
        oracle_lib.build_oracle(ref=os.path.isdir("/root/reference/src"))
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (oracle/_ref); built here when /root/reference exists, prebuilt on the GPU box."""
    import oracle_lib

    if not oracle_lib.Reference.available("strict"):
        if os.path.isdir("/root/reference/src"):
            oracle_lib.build_oracle(ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle_lib.Reference("strict")


@pytest.fixture(scope="session")
def engine():
    """The product library through its C ABI.  Fails (not skips) if it is missing or CUDA is unusable."""
    import mcmc_b200

    mcmc_b200.api.load()
    assert mcmc_b200.api.device_count() > 0, "no CUDA device visible to libmcmc_b200.so"
    return mcmc_b200
