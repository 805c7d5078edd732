import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # The shared libraries are build artefacts (git-ignored): in a fresh checkout compile them once, exactly as
    # __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU), so that the suite never runs
    # against a missing or silently absent product library.
    pkg = os.path.join(ROOT, "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200")
    if not os.path.exists(os.path.join(pkg, "libb200ret.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(pkg, "csrc")], stdout=subprocess.DEVNULL)
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    # `-m gpu` runs must fail loudly without a device rather than skip silently; plain runs
    # on a CPU-only box skip the gpu-marked tests.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
