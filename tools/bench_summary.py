import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.1f img/s  %.2f ms/step  e2e %.1f img/s (%.2f ms)  launch: %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"].get("launch")))
r = d["roofline"]; rb = d["roofline_bwd"]
print("F1 fwd %.1f us (%.1f GB/s, %.1f%%)  bwd %.1f us (%.1f%%)" % (r["avg_launch_us"], r["achieved"], 100 * r["frac"], rb["avg_launch_us"], 100 * rb["frac"]))
for k, v in d.get("roofline_conv", {}).items():
    print("conv %-6s %6.1f TF/s (%.1f%%)  %.2f ms/step  %d launches/step" % (k, v["achieved"], 100 * v["frac"], v["ms_per_step"], v.get("launches_per_step", v["launches_timed"] // 3)))
print("calls", d["config"].get("conv_calls_per_step"), "clocks", d.get("clocks"))
if "cpu_baseline" in d: print("cpu", d["cpu_baseline"])
