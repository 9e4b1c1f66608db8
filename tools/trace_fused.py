"""clock64 trace of the conv1a-fused conv1b launch inside a detector run (authoring aid).  AIRFE_TRACE_FUSED=1 is set here so that only
that launch stamps the buffer.  Slots: 0 producers start the TMEM -> A-stage pass of the tile (conv1a accumulators ready, stage free),
1 MMA warp start, 2 A stage landed, 3 conv1b MMAs issued, 4/5 epilogue warp 2 start/end, 6 A stage handed to the MMA warp.
Needs AIRFE_FUSE1A=1 in the environment (the fusion is experimental and off by default)."""
import os, sys
os.environ["AIRFE_TRACE_FUSED"] = "1"
os.environ.setdefault("AIRFE_FUSE1A", "1")
import numpy as np
import torch
sys.path.insert(0, ".")
from airslam_b200 import capi
from oracle import synth

B = 16
imgs = np.stack([synth.stereo_pair(752, 480, 100 + i)[0] for i in range(B)])
ctx = capi.Context(max_batch=B, enable_lightglue=0, enable_plnet=0)
for _ in range(2):
    ctx.detect_batch(capi.NET_SUPERPOINT, imgs)
buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
capi.lib().airfe_debug_conv_trace(buf.data_ptr())
ctx.detect_batch(capi.NET_SUPERPOINT, imgs)
capi.lib().airfe_debug_conv_trace(None)
t = buf.cpu().view(64, 8)
t0 = int(t[0, 0])
print("tile  prod_start  mma_start   A_landed mma_issued epi0_start   epi0_end  prod_done")
for i in range(30):
    if int(t[i, 0]) == 0:
        break
    print("%4d " % i + " ".join("%10d" % (int(v) - t0) for v in t[i][:7]))
