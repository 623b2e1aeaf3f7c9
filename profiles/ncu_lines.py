"""Aggregate ncu warp-stall samples per CUDA source line.
usage: python profiles/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; per = collections.defaultdict(lambda: collections.Counter()); src = {}; cur = None; fn = None
for r in rows:
    if not r: continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] not in ("", None) and r[0] != "-":
        cur = (fn, r[0]); src[cur] = r[1]
    d = dict(zip(hdr[4:], r[4:]))
    try: n = int(d["Warp Stall Sampling (All Samples)"])
    except Exception: continue
    c = per[cur]; c["_all"] += n; c["_inst"] += int(d.get("Instructions Executed") or 0)
    for k, v in d.items():
        if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0"): c[k] += int(v)
tot = sum(c["_all"] for c in per.values())
print("total samples", tot)
agg = collections.Counter()
for c in per.values():
    for k, v in c.items():
        if k.startswith("stall_"): agg[k] += v
print("by reason:", [(k, v) for k, v in agg.most_common(8)])
for key, c in sorted(per.items(), key=lambda kv: -kv[1]["_all"])[:topn]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(6) if k.startswith("stall_"))
    if key is None: continue
    print(f"{100*c['_all']/tot:5.1f}% inst={c['_inst']:>9} L{key[1]:>4} {src.get(key,'').strip()[:78]:78} | {st}")
