"""SASS census of the in-tree libairfe.so: per kernel, how many tcgen05 / TMA / TMEM instructions it contains (authoring aid; the committed
copy is profiles/r02_sass_census.txt).  The PTX names never appear in SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM / STTM,
cp.async.bulk.tensor -> UTMALDG / UTMASTG, legacy mma.sync -> HMMA (must be 0).

  python tools/sass_census.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "airslam_b200", "libairfe.so")
PATS = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
                                ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("STG.256", r"\bSTG\.E\.(ENL2\.)?256|STG\.[A-Z.]*256"), ("MUFU.EX2", r"MUFU\.EX2"),
                                ("SHFL", r"\bSHFL"), ("LDL/STL", r"\b(LDL|STL)\b")])


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        per[cur]["_insts"] += 1
        for k, p in PATS.items():
            if re.search(p, line):
                per[cur][k] += 1
    dem = {}
    try:
        names = "\n".join(per.keys())
        d = subprocess.run(["cu++filt"], input=names, capture_output=True, text=True).stdout.splitlines()
        dem = dict(zip(per.keys(), d))
    except Exception:
        pass
    lines = ["# SASS census of airslam_b200/libairfe.so (cuobjdump -sass, sm_100a): instruction counts per kernel",
             "# %-88s %7s " % ("kernel", "insts") + " ".join("%8s" % k for k in PATS)]
    tot = collections.Counter()
    for f, c in sorted(per.items(), key=lambda kv: -kv[1]["UTC*MMA"]):
        nm = dem.get(f, f)
        nm = nm[:nm.rfind("(")] if nm.rfind("(") > 0 else nm          # drop the parameter list, keep the template arguments
        nm = nm.replace("airfe::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
        lines.append("%-90s %7d " % (nm[:90], c["_insts"]) + " ".join("%8d" % c[k] for k in PATS))
        tot.update(c)
    lines.append("%-90s %7d " % ("TOTAL (%d kernels)" % len(per), tot["_insts"]) + " ".join("%8d" % tot[k] for k in PATS))
    txt = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
