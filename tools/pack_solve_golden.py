"""Compact golden of a whole solve (mocc_b200_solve .arrays of the REFERENCE sweepers): k history, every
`stride`-th flux entry, flux sum, pin powers.   python tools/pack_solve_golden.py <in.arrays> <out.arrays.gz> [stride]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocc_b200 import load_arrays, save_arrays  # noqa: E402

a = load_arrays(sys.argv[1])
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 97
flux = a["flux"]
save_arrays(sys.argv[2], {"k_history": a["k_history"], "flux_sample": np.ascontiguousarray(flux.reshape(-1)[::stride]),
                          "flux_stride": np.array([stride], dtype=np.int32),
                          "flux_shape": np.array(flux.shape, dtype=np.int64), "flux_sum": np.array([flux.sum()]),
                          "pin_powers": a["pin_powers"]})
print("wrote", sys.argv[2], "k", a["k_history"][-1], "outers", a["k_history"].size)
