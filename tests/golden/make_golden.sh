#!/bin/bash
# Regenerates the golden fixtures in this directory by running the UNMODIFIED
# reference (oracle/_ref/ref_tool, built by `make -C oracle ref` from the sources
# under /root/reference). Single-threaded so that reductions are deterministic.
#   *.mocflat.gz  flattened ray data of the case (mocc_b200/host/flatten.cpp)
#   *.golden.gz   inputs/outputs of selected reference sweep1g calls + k history
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
TOOL="$ROOT/oracle/_ref/ref_tool"
WORK="$(mktemp -d)"
trap 'rm -rf "$WORK"' EXIT
cp "$HERE"/inputs/* "$WORK"/
cp "$ROOT"/oracle/_ref/inputs/3x3.xml "$ROOT"/oracle/_ref/inputs/c5g7.xsl "$WORK"/
cd "$WORK"
export OMP_NUM_THREADS=1

run() { # name xml records [extra args...]
    local name="$1" xml="$2" rec="$3"; shift 3
    "$TOOL" golden "$xml" "$WORK/$name" --outers 2 --records "$rec" "$@" > "$WORK/$name.log" 2>&1 \
        || { tail -20 "$WORK/$name.log"; exit 1; }
    tail -1 "$WORK/$name.log"
    # geometry shared with mini2d_gs is stored once
    case "$name" in mini2d_jacobi|mini2d_nocmfd|mini3d_2d3d) ;; *) gzip -9 -n -c "$WORK/$name.mocflat" > "$HERE/$name.mocflat.gz";; esac
    gzip -9 -n -c "$WORK/$name.golden" > "$HERE/$name.golden.gz"
}

run mini2d_gs     mini2d.xml "0:0:0,0:1:2,1:2:2,1:0:1" --cmfd
run mini2d_jacobi mini2d.xml "0:0:0,0:1:2,1:2:2" --cmfd --set solver/sweeper@boundary_update=jacobi
run mini2d_nocmfd mini2d.xml "0:0:2,1:1:2"
run mini3d_gs     mini3d.xml "0:0:0,0:1:1,1:2:1" --cmfd
# MoCSweeper_2D3D + cmdo::CurrentCorrections (self-coupled): last inner records alpha/beta
run mini3d_2d3d   mini3d.xml "0:0:1,0:2:1,1:1:1" --2d3d
run 3x3_s05_gs    3x3.xml    "0:0:0,0:3:4,1:6:4" --cmfd --set solver/sweeper/rays@spacing=0.05
# geometry + cross sections of the reference's analytic test case (test_MoC_IHM.cpp): the oracle is run on it to the
# analytic infinite-medium spectrum in tests/test_oracle.py
run ihm           ihm.xml    "0:0:0"
ls -la "$HERE"/*.gz
# SECTIONS=records regenerates only the sweep1g / source records above
[ "${SECTIONS:-all}" = "records" ] && exit 0

# whole-solve goldens: the reference solver stack with the reference CPU sweepers
# (mocc_b200/bin/mocc_b200_solve = unmodified EigenSolver/CMFD/2D3D; <sweeper type> as in the input)
SOLVE="$ROOT/mocc_b200/bin/mocc_b200_solve"
python_pack() { python - "$1" "$2" <<'PY'
import sys
sys.path.insert(0, sys.argv[0] and ".")
from mocc_b200 import load_arrays, save_arrays
save_arrays(sys.argv[2], load_arrays(sys.argv[1]))
PY
}
for c in mini2d mini2d3d 3x3; do
    "$SOLVE" "$c.xml" "$WORK/$c.arrays" > "$WORK/$c.solve.log" 2>&1 || { tail -5 "$WORK/$c.solve.log"; exit 1; }
    tail -1 "$WORK/$c.solve.log"
    (cd "$ROOT" && python_pack "$WORK/$c.arrays" "$HERE/${c}_solve_ref.arrays.gz")
done

# C5G7 3-D (BASELINE.json config 4), FIRST OUTER ONLY of the 2D3D solve with the reference CPU sweepers, one thread.
# (From the second outer on the reference itself diverges on this input: profiles/r1/c5g7_3d.md.) Compact golden:
# k, every 97th flux entry, flux sum, pin powers.
python "$ROOT/tools/make_c5g7_3d.py" "$ROOT/mocc_b200/bin/inputs/c5g7_2d.xml" "$WORK/c5g7_3d.xml" --max-iter 1
cp "$ROOT/mocc_b200/bin/inputs/c5g7.xsl" "$WORK/"
(cd "$WORK" && OMP_NUM_THREADS=1 "$SOLVE" c5g7_3d.xml "$WORK/c5g7_3d.arrays" > "$WORK/c5g7_3d.solve.log" 2>&1) || { tail -5 "$WORK/c5g7_3d.solve.log"; exit 1; }
(cd "$ROOT" && python - "$WORK/c5g7_3d.arrays" "$HERE/c5g7_3d_outer1_ref.arrays.gz" <<'PY'
import sys
import numpy as np
from mocc_b200 import load_arrays, save_arrays
a = load_arrays(sys.argv[1])
flux = a["flux"]
save_arrays(sys.argv[2], {"k_history": a["k_history"], "flux_sample": np.ascontiguousarray(flux.reshape(-1)[::97]),
                          "flux_stride": np.array([97], dtype=np.int32), "flux_shape": np.array(flux.shape, dtype=np.int64),
                          "flux_sum": np.array([flux.sum()]), "pin_powers": a["pin_powers"]})
PY
)

# C5G7 3-D (BASELINE.json config 4) with the settings under which the reference's own 2D3D iteration settles
# (profiles/r2/c5g7_3d.md): transverse-leakage splitting on the MoC side, 10 Sn inners, one host thread.
#   c5g7_3d_12_ref : the full problem (3 x 3 assemblies of 17 x 17 pins, 9 planes), 12 outers: k stationary to
#                    +-3 pcm from outer 8 on (the reference's iteration breaks down at outer 14)
#   c5g7_3d_n9_ref : the same problem with every assembly cut to its central 9 x 9 pins (ray spacing 0.045), 24 outers: k stationary
#                    to +-1.5 pcm from outer 9 on
# Compact goldens (tools/pack_solve_golden.py): k history, every 97th flux entry, flux sum, pin powers.
for v in "12:--max-iter 12" "n9:--lattice-n 9 --spacing 0.045 --max-iter 24"; do
    name="c5g7_3d_${v%%:*}"
    python "$ROOT/tools/make_c5g7_3d.py" "$ROOT/mocc_b200/bin/inputs/c5g7_2d.xml" "$WORK/$name.xml" ${v#*:} \
        --sn-inner 10 --moc-attrs 'tl_splitting="t"'
    (cd "$WORK" && OMP_NUM_THREADS=1 "$SOLVE" $name.xml "$WORK/$name.arrays" > "$WORK/$name.solve.log" 2>&1) \
        || { tail -5 "$WORK/$name.solve.log"; exit 1; }
    python "$ROOT/tools/pack_solve_golden.py" "$WORK/$name.arrays" "$HERE/${name}_ref.arrays.gz"
done
