"""CPU-side checks of the drop-in boundary: the C-ABI library is built in-tree, loads,
and exports every symbol include/mocc_b200.h declares. No compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_case


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mocc_b200.h")).read()
    return sorted(set(re.findall(r"\b(mocb200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from mocc_b200 import LIB_PATH
    assert os.path.exists(LIB_PATH), "build with __graft_entry__.build()"
    lib = ctypes.CDLL(LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mocc_b200.h but not exported"


def test_version_string():
    from mocc_b200 import load_library
    assert b"mocc_b200" in load_library().mocb200_version()


def test_problem_struct_layout_matches_header():
    """ctypes mirror of struct mocb200_problem follows the header field order."""
    from mocc_b200.capi import _SCALARS, _ARRAYS
    hdr = open(os.path.join(ROOT, "include", "mocc_b200.h")).read()
    body = hdr[hdr.index("typedef struct mocb200_problem {"):hdr.index("} mocb200_problem;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.split("{")[-1].strip()
        if not decl:
            continue
        parts = decl.replace("*", " ").split(",")
        first = parts[0].split()
        names.append(first[-1])
        names.extend(p.strip() for p in parts[1:])
    assert names == [n for n, _ in _SCALARS + _ARRAYS]


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mocc_b200 import Sweeper
    flat, _ = load_case("mini2d_gs")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        Sweeper(flat)


def test_flatfile_roundtrip(tmp_path):
    from mocc_b200.flatfile import load_arrays, save_arrays
    flat, _ = load_case("mini2d_gs")
    p = tmp_path / "x.mocflat.gz"
    save_arrays(p, flat)
    back = load_arrays(p)
    assert list(back) == list(flat)
    for k in flat:
        assert back[k].dtype == flat[k].dtype and np.array_equal(back[k], flat[k])


def test_material_tables_reproduce_the_per_fsr_cross_sections():
    """capi.material_tables: cross-section-mesh regions out of per-FSR arrays (the form mocb200_set_source_xs takes)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from conftest import load_case
    from mocc_b200.capi import material_tables
    flat, gold = load_case("3x3_s05_gs")
    G = int(flat["n_group"][0])
    xs_nf = np.stack([gold[f"xs_nf_{g}"] for g in range(G)])
    xs_ch = np.stack([gold[f"xs_ch_{g}"] for g in range(G)])
    xs_scat = np.stack([gold[f"xs_scat_to_{g}"].reshape(G, -1) for g in range(G)])
    fsr_mat, t_nf, t_ch, t_scat = material_tables(xs_nf, xs_ch, xs_scat)
    assert fsr_mat.dtype == np.int32 and fsr_mat.size == xs_nf.shape[1] and 1 < t_nf.shape[0] <= 8
    assert np.array_equal(t_nf[fsr_mat].T, xs_nf) and np.array_equal(t_ch[fsr_mat].T, xs_ch)
    assert np.array_equal(np.moveaxis(t_scat[fsr_mat], 0, 2), xs_scat)
