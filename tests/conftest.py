import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# case name -> (flat file, golden file)
CASES = {
    "mini2d_gs": ("mini2d_gs.mocflat.gz", "mini2d_gs.golden.gz"),
    "mini2d_jacobi": ("mini2d_gs.mocflat.gz", "mini2d_jacobi.golden.gz"),
    "mini2d_nocmfd": ("mini2d_gs.mocflat.gz", "mini2d_nocmfd.golden.gz"),
    "mini3d_gs": ("mini3d_gs.mocflat.gz", "mini3d_gs.golden.gz"),
    "3x3_s05_gs": ("3x3_s05_gs.mocflat.gz", "3x3_s05_gs.golden.gz"),
    "mini3d_2d3d": ("mini3d_gs.mocflat.gz", "mini3d_2d3d.golden.gz"),
    # geometry and cross sections of the reference's analytic test (test_MoC_IHM.cpp)
    "ihm": ("ihm.mocflat.gz", "ihm.golden.gz"),
}

_cache = {}


def load_case(name):
    from mocc_b200.flatfile import load_arrays
    if name not in _cache:
        flat, gold = CASES[name]
        _cache[name] = (load_arrays(os.path.join(GOLDEN, flat)), load_arrays(os.path.join(GOLDEN, gold)))
    return _cache[name]


def records(gold):
    n = int(gold["n_rec"][0])
    out = []
    for r in range(n):
        p = f"rec{r}_"
        out.append({k[len(p):]: v for k, v in gold.items() if k.startswith(p)})
    return out
