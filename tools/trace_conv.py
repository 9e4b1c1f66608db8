"""clock64 trace of CTA 0's first tiles in the halo-reuse conv kernel (authoring aid; see airfe_debug_conv_trace).
usage: trace_conv.py [c64|c64po|c6432|c32|c128]   prints per-tile stage stamps in cycles relative to the first producer issue."""
import sys
import torch
sys.path.insert(0, ".")
from airslam_b200 import capi

lib = capi.lib()
sys.argv = sys.argv[:1] + (sys.argv[1:] or ["c64po"])
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
import tools.prof_conv as pc   # noqa: E402  (runs the listed configs once untraced = warm-up)

for tag in sys.argv[1:]:
    cfg = {"c64": (64, 64, True, True, 512), "c64np": (64, 64, True, False, 512), "c64po": (64, 64, False, True, 512), "c6432": (64, 32, True, False, 512),
           "c32": (32, 32, False, True, 512), "c128": (128, 128, True, False, 256)}[tag]
    cin, cout, full, pool, hw = cfg
    buf.zero_()
    lib.airfe_debug_conv_trace(buf.data_ptr())
    pc.conv(tag + "-traced", 16, hw, hw, cin, cout, full, pool, reps=2)
    lib.airfe_debug_conv_trace(None)
    t = buf.cpu().view(64, 8)
    t0 = int(t[0, 0])
    print("tile   prod_issue  mma_start  A_landed  mma_issued  epi0_start epi0_end  epi7_start epi7_end   (cycles since first TMA issue)")
    for i in range(0, 40):
        if int(t[i, 0]) == 0:
            break
        print("%4d " % i + " ".join("%10d" % (int(v) - t0) for v in t[i]))
    d = [(int(t[i + 1, 3]) - int(t[i, 3])) for i in range(8, 38) if int(t[i + 1, 3])]
    if d:
        print("steady-state cycles per tile (MMA issue to MMA issue): mean %.0f" % (sum(d) / len(d)))
