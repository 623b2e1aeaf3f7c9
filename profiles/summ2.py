import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d["roofline"]
        print("N", d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "frac", r["frac"], "launch_ms", r["launch_ms"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
        print("   parity", d["parity"])
    elif "rror" in l or "Traceback" in l:
        print(l.rstrip())
