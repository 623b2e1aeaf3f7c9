"""Histogram of the Blackwell-specific SASS opcodes in the built library, per kernel family:
cuobjdump -sass jets.jl_b200/libjets_b200.so -> profiles/sass_opcodes.txt.  Runs on the CPU box (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jets.jl_b200", "libjets_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "LDT", "STT", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "ACQBULK",
         "LDG.E.128", "STG.E.128", "LD.E.STRONG.SYS", "ST.E.STRONG.SYS", "LDG.E.STRONG.SYS", "STG.E.STRONG.SYS", "MEMBAR.ALL.SYS", "MEMBAR.SC.SYS",
         "NANOSLEEP", "FFMA", "DFMA", "HMMA", "BAR.SYNC", "ATOMG", "RED", "REDG", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "")
            name = re.sub(r"\(.*", "", name)
            name = re.sub(r"^void ", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur is not None:
            op = m.group(1)
            cur["__all__"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    cur[w] += 1
    lines = ["# SASS opcode histogram of jets.jl_b200/libjets_b200.so (sm_100a), by kernel; produced by profiles/sass_opcodes.py",
             "# columns: kernel | total instructions | watched opcodes (count)", ""]
    tot = collections.Counter()
    for name, c in per.items():
        hits = ", ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        lines.append(f"{name} | {c['__all__']} | {hits}")
        tot.update(c)
    lines += ["", "TOTAL | %d | %s" % (tot["__all__"], ", ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]))]
    dst = os.path.join(ROOT, "profiles", "sass_opcodes.txt")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(dst, len(per), "kernels")


if __name__ == "__main__":
    sys.exit(main())
