O=gpurun_out/final2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_plugin.py -m gpu -q -k "eigenvalue_solve or 2d3d_solve or rehomog or device_built or group_batched or jacobi" 2>&1 | tail -3 | tee $O/pytest_plugin_subset.txt
python bench.py > $O/bench_pergroup.json 2> $O/bench_pergroup.err; tail -c 300 $O/bench_pergroup.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/final2/bench_pergroup.json") if l.startswith("{")][-1])
r = d["roofline"]
print("value %.4g e2e %.4g ms/step %.3f frac %.4f ms/inner %.5f e2e_dev %.4g plugin %.4g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["ms_per_launch"], d["e2e_device_sources"]["value"], d["e2e_plugin"]["value"]), d["time_to_converge_s"], d["cpu_baseline"]["value"])
PY
