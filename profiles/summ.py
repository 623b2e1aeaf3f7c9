import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d["roofline"]
        print("C5 value", d["value"], "frac", r["frac"], "launch_ms", r["launch_ms"], "e2e", d["e2e"]["value"], "clk", d["clocks"])
        for k, v in d.get("other_workloads", {}).items():
            print(" ", k, {x: v[x] for x in v if x in ("value", "frac_of_hbm_peak", "ms_per_step", "forward_gbs", "adjoint_gbs", "error")})
        if "cpu_baseline" in d: print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("faithful_single_thread", {}).get("value"))
    else:
        print(l.rstrip())
