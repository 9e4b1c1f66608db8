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


def allk(path, out, note, hbm_gbs=6569.6, json_out=None, pairs=None):
    """Every kernel of the last captured chunk: per kernel name -> launches, time, share, DRAM bytes and GB/s (vs the measured HBM peak),
    time-weighted tensor-pipe %, L2 bytes; then the launch-by-launch list.  Also writes the conv traffic json bench.py reads."""
    _, rs = rows(path)
    by_id = {}
    for r in rs:
        d = by_id.setdefault(int(r["ID"]), {"name": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
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
        elif m.startswith("sm__warps_active"):
            d["occ"] = float(r["Metric Value"])
    ids = sorted(by_id)
    starts = [i for i in ids if by_id[i]["name"].startswith("resize_kernel") or by_id[i]["name"].startswith("lg_prepare") and not any(by_id[j]["name"].startswith("resize_kernel") for j in ids)]
    if len(starts) >= 2:
        ids = [i for i in ids if i >= starts[-1]]            # the capture holds the warm-up chunk and the timed chunk: keep the last one
    per = {}
    for i in ids:
        d = by_id[i]
        a = per.setdefault(d["name"], {"n": 0, "us": 0.0, "mb": 0.0, "tp_us": 0.0, "l2": 0.0})
        a["n"] += 1
        a["us"] += d.get("us", 0)
        a["mb"] += d.get("rd", 0) + d.get("wr", 0)
        a["tp_us"] += d.get("tp", 0) * d.get("us", 0)
        a["l2"] += d.get("l2", 0)
    tot = sum(a["us"] for a in per.values())
    with open(out, "w") as f:
        f.write("# %s\n" % note)
        f.write("# ncu --clock-control none: per-launch times are cold-cache and serialised (compare SHARES, not absolutes).  %d launches, %.1f us\n" % (len(ids), tot))
        f.write("# HBM GB/s = (dram read + write) / time; %% of the measured copy peak %.1f GB/s.  tensor%% = time-weighted sm__pipe_tensor_cycles_active\n" % hbm_gbs)
        f.write("%-46s %5s %10s %6s %10s %9s %7s %8s %10s\n" % ("kernel", "n", "time us", "share", "dram MB", "GB/s", "%peak", "tensor%", "L2 MB"))
        for k in sorted(per, key=lambda k: -per[k]["us"]):
            a = per[k]
            gbs = a["mb"] * 1e-3 / (a["us"] * 1e-6) if a["us"] else 0
            f.write("%-46s %5d %10.1f %5.1f%% %10.1f %9.1f %6.1f%% %7.1f%% %10.1f\n" % (k[:46], a["n"], a["us"], 100 * a["us"] / tot, a["mb"], gbs, 100 * gbs / hbm_gbs,
                                                                                    a["tp_us"] / a["us"] if a["us"] else 0, a["l2"]))
        f.write("\n# launch by launch: id | kernel | grid | us | dram MB | GB/s | tensor %% | L2 MB\n")
        for n, i in enumerate(ids):
            d = by_id[i]
            mb = d.get("rd", 0) + d.get("wr", 0)
            f.write("%3d %-40s %-16s %9.1f %9.2f %8.1f %6.1f%% %9.1f\n" % (n, d["name"][:40], d["grid"], d.get("us", 0), mb, mb * 1e-3 / (d.get("us", 1) * 1e-6) if d.get("us") else 0,
                                                                     d.get("tp", 0), d.get("l2", 0)))
    if json_out:
        cv = [by_id[i] for i in ids if by_id[i]["name"].startswith("tc_conv3x3")]
        tb = sum(d.get("rd", 0) + d.get("wr", 0) for d in cv) * 1e6
        json.dump({"pairs_per_step": pairs, "kernel": "tc_conv3x3_kernel", "launches": len(cv), "dram_bytes_per_step": tb, "dram_bytes_per_launch": tb / max(1, len(cv)),
                   "tensor_pipe_pct_time_weighted": sum(d.get("tp", 0) * d.get("us", 0) for d in cv) / max(1e-9, sum(d.get("us", 0) for d in cv)), "source": out}, open(json_out, "w"))
    print("wrote", out)


if __name__ == "__main__":
    if sys.argv[1] == "all":
        allk(sys.argv[2], sys.argv[3], sys.argv[4], json_out=sys.argv[5] if len(sys.argv) > 5 else None, pairs=int(sys.argv[6]) if len(sys.argv) > 6 else None)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        traffic(sys.argv[2], sys.argv[3], int(sys.argv[4]))
