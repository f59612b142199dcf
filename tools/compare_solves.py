"""Compare two mocc_b200_solve outputs: k-eff (pcm) and FSR scalar flux (max relative)."""
import sys, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocc_b200 import load_arrays
a, b = load_arrays(sys.argv[1]), load_arrays(sys.argv[2])
ka, kb = a["k_history"], b["k_history"]
fa, fb = a["flux"], b["flux"]
# eigenvector normalisation is fixed by the solver, but compare shape robustly too
rel = np.max(np.abs(fa - fb) / np.abs(fa))
scale = (fa * fb).sum() / (fb * fb).sum()
rel_n = np.max(np.abs(fa - scale * fb) / np.abs(fa))
print(json.dumps({"a": sys.argv[1], "b": sys.argv[2], "k_a": float(ka[-1]), "k_b": float(kb[-1]),
                  "dk_pcm": float((kb[-1] - ka[-1]) * 1e5), "outers_a": int(ka.size), "outers_b": int(kb.size),
                  "flux_max_rel": float(rel), "flux_max_rel_renormalised": float(rel_n),
                  "sweep_s_a": float(a["sweep_seconds"][0]), "sweep_s_b": float(b["sweep_seconds"][0]),
                  "solve_s_a": float(a["solve_seconds"][0]), "solve_s_b": float(b["solve_seconds"][0]),
                  "device_sweep_ms_b": float(b["device_sweep_ms"][0])}))
