"""Turn ncu CSV logs of one bench.py device step into the summaries kept under profiles/ (authoring aid; nothing imports this).

  launch list :  ncu --metrics gpu__time_duration.sum --clock-control none -s S -c C --csv --log-file gpurun_out/launches.csv \
                     python bench.py --steps 1 --warmup 1 --device-only
                 python tools/ncu_summarize.py launches gpurun_out/launches.csv profiles/r01_ncu_launch_list.txt "<header note>"
  conv traffic:  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
                     sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none \
                     -k regex:tc_conv3x3 -s N -c N --csv --log-file gpurun_out/conv_traffic.csv python bench.py --steps 1 --warmup 1 --device-only
                 python tools/ncu_summarize.py traffic gpurun_out/conv_traffic.csv profiles/r01_ncu_conv3x3_traffic.txt PAIRS
                 (also writes profiles/r01_conv3x3_traffic.json, which bench.py reads for roofline.traffic)
"""
import csv
import json
import re
import sys


def rows(path):
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    rd = list(csv.reader(lines))
    hdr = rd[0]
    return hdr, [dict(zip(hdr, r)) for r in rd[1:] if len(r) == len(hdr)]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    u = unit.strip().lower()
    return v * (1e-3 if u.startswith("n") else 1.0 if u.startswith("u") else 1e3 if u.startswith("m") else 1e6)


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("airfe::", "")


def launches(path, out, note):
    _, rs = rows(path)
    seq = [(short(r["Kernel Name"]), to_us(r["Metric Value"], r["Metric Unit"])) for r in rs if r["Metric Name"] == "gpu__time_duration.sum"]
    # the capture holds the warm-up step(s) and the timed step: keep the LAST complete step (a step starts with resize_kernel)
    starts = [i for i, (k, _) in enumerate(seq) if k.startswith("resize_kernel")]
    if len(starts) >= 2:
        seq = seq[starts[-1]:] if len(seq) - starts[-1] >= starts[-1] - starts[-2] else seq[starts[-2]:starts[-1]]
    per = {}
    for k, us in seq:
        per.setdefault(k, [0, 0.0])
        per[k][0] += 1
        per[k][1] += us
    tot = sum(v[1] for v in per.values())
    n = sum(v[0] for v in per.values())
    with open(out, "w") as f:
        f.write("# %s\n" % note)
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  total %.1f us over %d launches\n" % (tot, n))
        for k in sorted(per, key=lambda k: -per[k][1]):
            f.write("%-70s launches %4d  time %10.1f us  share %5.1f%%\n" % (k[:70], per[k][0], per[k][1], 100 * per[k][1] / tot))
    print("wrote", out, "(%d launches, %.1f us)" % (n, tot))


def traffic(path, out, pairs):
    _, rs = rows(path)
    by_id = {}
    for r in rs:
        d = by_id.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"])})
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = to_us(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_read.sum":
            d["rd"] = to_mb(r["Metric Value"], r["Metric Unit"])
        elif m == "dram__bytes_write.sum":
            d["wr"] = to_mb(r["Metric Value"], r["Metric Unit"])
        elif m.startswith("sm__pipe_tensor_cycles_active"):
            d["tp"] = float(r["Metric Value"])
        elif m == "lts__t_bytes.sum":
            d["l2"] = to_mb(r["Metric Value"], r["Metric Unit"])
    ids = sorted(by_id)
    if len(ids) % 2 == 0 and len(ids) > 80:
        ids = ids[len(ids) // 2:]          # warm-up step + timed step captured: keep the timed one
    tot_b = sum((by_id[i].get("rd", 0) + by_id[i].get("wr", 0)) for i in ids) * 1e6
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active...,lts__t_bytes.sum "
                "--clock-control none -k regex:tc_conv3x3 -s %d -c %d\n" % (len(ids), len(ids)))
        f.write("# all %d tc_conv3x3_kernel launches of ONE device step (P = %d stereo pairs = %d frames, PLNet).  id: us | dram read MB | dram write MB | tensor pipe %% | L2 MB\n"
                % (len(ids), pairs, 2 * pairs))
        for n, i in enumerate(ids):
            d = by_id[i]
            f.write("%3d: %9.1f us %10.2f %10.2f %7.1f%% %10.1f\n" % (n, d.get("us", 0), d.get("rd", 0), d.get("wr", 0), d.get("tp", 0), d.get("l2", 0)))
        f.write("# total dram bytes per step %.3f GB, per launch %.1f MB\n" % (tot_b / 1e9, tot_b / 1e6 / max(1, len(ids))))
    js = {"pairs_per_step": pairs, "kernel": "tc_conv3x3_kernel", "launches": len(ids), "dram_bytes_per_step": tot_b,
          "dram_bytes_per_launch": tot_b / max(1, len(ids)), "source": out}
    jp = out.replace("r01_ncu_conv3x3_traffic.txt", "r01_conv3x3_traffic.json")
    json.dump(js, open(jp, "w"))
    print("wrote", out, jp)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]))
