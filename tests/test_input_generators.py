"""The authored inputs of BASELINE.json configs 4 and 5 (tools/make_c5g7_3d.py, tools/make_quarter_core.py) are
derived from the reference's 2-D example at test time; check their structure (no GPU, no solver run)."""
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "mocc_b200", "bin", "inputs", "c5g7_2d.xml")
pytestmark = pytest.mark.skipif(not os.path.exists(SRC), reason="example inputs not staged (build with the reference)")


def _parse(path):
    return ET.fromstring("<root>" + open(path).read() + "</root>")  # MOCC inputs have several top-level nodes


def test_c5g7_3d_input(tmp_path):
    out = tmp_path / "c5g7_3d.xml"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_c5g7_3d.py"), SRC, str(out), "--max-iter", "1"],
                          stdout=subprocess.DEVNULL)
    r = _parse(out)
    sw = r.find("solver/sweeper")
    assert sw.get("type") == "2d3d" and r.find("solver").get("max_iter") == "1" and r.find("solver").get("cmfd") == "t"
    assert sw.find("moc_sweeper/rays").get("spacing") == "0.05" and sw.find("sn_sweeper").get("equation") == "cdd"
    for asm in r.findall("assembly"):
        lats = asm.find("lattices").text.split()
        hz = [float(x) for x in asm.find("hz").text.split()]
        assert int(asm.get("np")) == len(lats) == len(hz) == 9
        assert lats[:3] == ["3", "3", "3"]                       # axial reflector on top (read top-down)
        assert abs(sum(hz[3:]) - 42.84) < 1e-9 and abs(sum(hz[:3]) - 21.42) < 1e-9
    core = r.find("core")
    assert core.get("top") == "vacuum" and core.get("bottom") == "reflect"


def test_quarter_core_input(tmp_path):
    out = tmp_path / "qc.xml"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_quarter_core.py"), SRC, str(out), "--n", "5"],
                          stdout=subprocess.DEVNULL)
    core = _parse(out).find("core")
    rows = [ln.split() for ln in core.text.strip().splitlines()]
    assert core.get("nx") == core.get("ny") == "5" and len(rows) == 5 and all(len(x) == 5 for x in rows)
    assert all(x[-1] == "3" for x in rows) and rows[-1] == ["3"] * 5      # reflector ring east / south
    assert rows[0][:4] == ["1", "2", "1", "2"] and rows[1][:4] == ["2", "1", "2", "1"]
