import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def built():
    """The CUDA library and the oracle must be built (driver: __graft_entry__.build())."""
    import __graft_entry__ as g
    from vk_raytrace_b200.lib import library_path
    from oracle import binding as ob
    if not os.path.exists(library_path()):
        g.build()
    ob.build()
    return True


@pytest.fixture(scope="session")
def ob(built):
    from oracle import binding
    binding.load()
    return binding


@pytest.fixture(scope="session")
def vk(built):
    import vk_raytrace_b200
    vk_raytrace_b200.load_library()
    return vk_raytrace_b200


@pytest.fixture(scope="session")
def gpu(vk):
    ctx = vk.Context(0)
    # oracle comparisons use the IEEE-exact shade kernel (vkrt_render_opts.precise_shading = 1); the default fast shade
    # kernel has its own tests (test_fast_shading_*) and is what the full-size property tests and bench.py run
    ctx.precise_default = True
    yield ctx
    ctx.close()


def rel_l2(a, b):
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))
