"""Print the per-kernel breakdown of bench.py JSON lines (files given on the command line)."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "ms/step %.3f" % d["ms_per_step"], "value %.4g" % d["value"],
          "e2e ms %s" % d.get("e2e", {}).get("ms_per_step"), "launches", d.get("gpu_launches"))
    r = d.get("roofline", {})
    print("   roofline:", {k: r.get(k) for k in ("kernel", "achieved", "frac", "traffic", "step_alg_gbs")},
          "streaming:", r.get("streaming_kernels"))
    for k in d.get("kernels", []):
        print("   %-20s %8.3f ms  x%-4s %7.1f GB/s" % (k["name"], k["ms_per_step"],
                                                       k["launches_per_step"], k["gbs"]))
    res = d.get("result", {})
    print("   phase_ms", res.get("phase_ms"), "kept", res.get("n_kept"), "partial",
          res.get("partial_bundles"), res.get("partial_candidates"))
