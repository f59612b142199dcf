"""The flattener (mocc_b200/host/flatten.cpp: append_ray) against the known answers of the reference's own ray
tests: src/sweepers/moc/tests/test_Ray.cpp:58-91 ("simple_ray": coarse-mesh linkage of two rays through corners)
and :270-305 ("weird_ray": the 36-segment golden vector). tests/golden/flatten_rays.json is what
oracle/_ref/flatten_ray_check printed (reference Ray objects pushed through append_ray); when that tool and the
reference's test inputs are present (the build container) it is re-run and must reproduce the fixture bit for bit.
FSR indexing and ray linkage must be exact; lengths to the tolerance test_Ray.cpp itself uses.
"""
import json
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIXTURE = os.path.join(HERE, "golden", "flatten_rays.json")
TOOL = os.path.join(ROOT, "oracle", "_ref", "flatten_ray_check")
REF_TESTS = "/root/reference/src/sweepers/moc/tests"
EAST, NORTH, WEST, SOUTH = 0, 1, 2, 3  # Surface enum, src/core/constants.hpp:36-39


def unpack(c):
    """RayCoarseData {fw:4, bw:4, nseg_fw:8, nseg_bw:8} (ray.hpp:42-46) as flatten.cpp packs it"""
    return c & 0xF, (c >> 4) & 0xF, (c >> 8) & 0xFF, (c >> 16) & 0xFF


def track(d, t):
    s0, s1 = d["trk_seg_begin"][t], d["trk_seg_begin"][t + 1]
    c0, c1 = d["trk_cm_begin"][t], d["trk_cm_begin"][t + 1]
    cell_fw, cell_bw, surf_fw, surf_bw = d["trk_cm_start"][4 * t: 4 * t + 4]
    return {"bc": d["trk_bc"][2 * t: 2 * t + 2], "cell_fw": cell_fw, "cell_bw": cell_bw, "surf_fw": surf_fw,
            "surf_bw": surf_bw, "fsr": d["seg_fsr"][s0:s1], "len": d["seg_len"][s0:s1],
            "cm": [unpack(c) for c in d["cm_data"][c0:c1]]}


def test_simple_ray_coarse_linkage():
    d = json.load(open(FIXTURE))["simple_ray"]
    r = track(d, 0)  # Ray((0, 1) -> (4, 5)): starts on a corner, ends on a corner, crosses corners (test_Ray.cpp:58-91)
    assert (r["surf_fw"], r["cell_fw"], r["surf_bw"], r["cell_bw"]) == (37, 6, 88, 27)
    assert len(r["fsr"]) == 12 and len(r["cm"]) == 8
    assert np.allclose(r["len"], math.sqrt(2.0) / 3.0, rtol=0, atol=1e-5)
    assert [c[0] for c in r["cm"]] == [EAST, NORTH] * 4
    assert [c[1] for c in r["cm"]] == [WEST, SOUTH, WEST, SOUTH, WEST, SOUTH, SOUTH, WEST]
    assert [c[2] for c in r["cm"]] == [3, 0] * 4 and [c[3] for c in r["cm"]] == [3, 0] * 4
    r = track(d, 1)  # Ray((4, 0) -> (6, 2)) (test_Ray.cpp:93-103)
    assert (r["surf_fw"], r["cell_fw"], r["surf_bw"], r["cell_bw"]) == (89, 4, 43, 11)
    assert len(r["fsr"]) == 6 and len(r["cm"]) == 4


def test_weird_ray_golden_segments():
    r = track(json.load(open(FIXTURE))["weird_ray"], 0)
    seg_index_expect = [29, 28, 27, 26, 25, 105, 104, 103, 102, 101, 100, 180, 179, 178, 177, 176, 175, 255,
                        254, 253, 252, 251, 250, 330, 329, 328, 327, 326, 325, 405, 404, 403, 402, 401, 400, 480]
    seg_len_expect = [
        0.12752525252525659, 0.12752525252525249, 0.12752525252525249, 0.12752525252525268, 0.12114898989898612,
        0.0063762626262663788, 0.12752525252525249, 0.12752525252525249, 0.12752525252525249, 0.12752525252525249,
        0.11477272727272445, 0.012752525252528228, 0.12752525252525249, 0.12752525252525249, 0.12752525252525249,
        0.12752525252525249, 0.10839646464646216, 0.01912878787879034, 0.12752525252525268, 0.12752525252525249,
        0.12752525252525249, 0.12752525252525249, 0.10202020202019987, 0.025505050505052623, 0.12752525252525268,
        0.12752525252525249, 0.12752525252525249, 0.12752525252525249, 0.095643939393937588, 0.031881313131314912,
        0.12752525252525249, 0.12752525252525249, 0.12752525252525268, 0.12752525252525249, 0.089267676767675302,
        0.038257575757577197]
    assert r["bc"] == [106, 7]
    assert r["fsr"] == seg_index_expect  # exact (test_Ray.cpp:281-288, 303)
    assert np.allclose(r["len"], seg_len_expect, rtol=0, atol=1e-15)  # test_Ray.cpp:304


def test_fixture_is_what_the_reference_produces(tmp_path):
    if not (os.path.exists(TOOL) and os.path.isdir(REF_TESTS)):
        pytest.skip("oracle/_ref/flatten_ray_check or the reference's test inputs are not here (GPU box)")
    out = tmp_path / "rays.json"
    subprocess.run([TOOL, REF_TESTS, str(out)], cwd=os.path.join(ROOT, "oracle", "_ref", "inputs"), check=True,
                   stdout=subprocess.DEVNULL)
    assert json.load(open(out)) == json.load(open(FIXTURE))


@pytest.mark.parametrize("case", ["mini2d_gs", "mini3d_gs", "3x3_s05_gs", "ihm"])
def test_flattened_rays_preserve_every_fsr_volume(case):
    """The reference's test_RayData (src/sweepers/moc/tests/test_RayData.cpp:50-71) on the FLATTENED arrays: summed over
    the sweep angles of octants 1-2, segment length x ray spacing x angle weight x 2 pi gives 4 pi x the FSR's area,
    for every FSR of every macroplane (the reference checks 1e-14 absolute on its square; relative 1e-13 here, the
    cases differ in size). With the flat volume correction this even holds angle by angle."""
    import numpy as np
    from conftest import load_case
    flat, _ = load_case(case)
    n_ang, n_geom, n_reg = (int(flat[k][0]) for k in ("n_ang", "n_geom", "n_reg"))
    gtb, tsb = flat["geom_trk_begin"], flat["trk_seg_begin"]
    first = list(flat["plane_first_reg"]) + [n_reg]
    # vol carries the plane height (Mesh volumes); the rays see areas
    height = flat["plane_height"]
    for ip, u in enumerate(flat["plane_unique"]):
        lo, hi = first[ip], first[ip + 1]
        area = flat["vol"][lo:hi] / height[ip]
        total = np.zeros(hi - lo)
        for a in range(n_ang):
            g = int(u) * n_geom + int(flat["ang_geom"][a])
            s0, s1 = int(tsb[gtb[g]]), int(tsb[gtb[g + 1]])
            per_angle = np.bincount(flat["seg_fsr"][s0:s1], weights=flat["seg_len"][s0:s1], minlength=hi - lo) * flat["ang_spacing"][a]
            assert np.max(np.abs(per_angle - area) / area) < 1e-12, f"plane {ip} angle {a}"
            total += per_angle * flat["ang_weight"][a] * 2.0 * np.pi
        assert np.max(np.abs(total - 4.0 * np.pi * area) / (4.0 * np.pi * area)) < 1e-13
