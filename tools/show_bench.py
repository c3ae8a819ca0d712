"""Print the headline and the per-kernel table of a bench.py JSON line (dev aid)."""
import json
import sys

for f in sys.argv[1:]:
    t = open(f).read().strip()
    if not t:
        print(f, "EMPTY")
        continue
    d = json.loads(t.split("\n")[-1])
    print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), "seq",
          d["config"].get("sequential_ms_per_step"), "launches", d["gpu_launches"], "clocks", d.get("clocks"))
    r = d["roofline"]
    print(" roofline", {k: r.get(k) for k in ("kernel", "achieved", "frac", "launches_per_step", "avg_launch_ms", "share_of_step",
                                              "hbm_GBps", "mma_issued_frac")})
    print(" stages", d.get("stage_us_per_scene"))
    tot = 0.0
    for k in d["kernels"]:
        tot += k["ms_per_step"]
        print("   %-62s x%-3g %7.4f %7.4f  GB/s %-8s TF %s" % (k["kernel"], k["calls_per_step"], k["avg_ms"], k["ms_per_step"],
                                                              k["algorithmic_GBps"], k["algorithmic_TFLOPs"]))
    print("   sum of listed kernels: %.3f ms" % tot)
