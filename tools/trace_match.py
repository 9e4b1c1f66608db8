"""clock64 trace of CTA 0 of the fused LightGlue kernels (authoring aid; see airfe_debug_match_trace).
usage: trace_match.py [pairs]   -- matches `pairs` pairs of 400-keypoint sets (default 47 = the bench chunk) and prints, for the LAST
tc_ffn / tc_attn launch, per-tile stage stamps in cycles relative to the tile's first stamp."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from airslam_b200 import capi
from oracle import synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 47
lib = capi.lib()
f0 = [synth.keypoint_set(400, 752, 480, 70 + k) for k in range(P)]
f1 = [synth.keypoint_set(400, 752, 480, 170 + k, perturb_of=f0[k])[0] for k in range(P)]
ctx = capi.Context(max_batch=P, enable_superpoint=1, enable_plnet=0)
for _ in range(2):
    ctx.match_batch(capi.MATCHER_LIGHTGLUE, f0, f1)
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
lib.airfe_debug_match_trace(buf.data_ptr())
t = time.time()
ctx.match_batch(capi.MATCHER_LIGHTGLUE, f0, f1)
lib.airfe_debug_match_trace(None)
torch.cuda.synchronize()
b = buf.cpu().numpy()
ffn = b[:128].reshape(8, 16)
att = b[1024:1152].reshape(8, 16)
print("tc_ffn, CTA 0 (cycles; MMA warp: p1_ready p1_issued p2_ready p2_issued p3_ready p3_issued | epilogue warp 2: acc1 epi1_done acc2 ln_mean ln_var epi2_done acc3 epi3_done)")
for i in range(8):
    if ffn[i, 0] == 0:
        break
    t0 = ffn[i, 0]
    print("tile %d  start@%d " % (i, t0 - ffn[0, 0]) + " ".join("%7d" % (v - t0) for v in ffn[i, :14]))
print("tc_attn, CTA 0 (cycles; MMA warp: reach q_ready pv_issued _ | softmax warp 2: s_full max_done exp_done o_full ctx_stored)")
for i in range(8):
    if att[i, 0] == 0:
        break
    t0 = att[i, 0]
    print("tile %d  start@%d " % (i, t0 - att[0, 0]) + " ".join("%7d" % (v - t0) for v in att[i, :9]))
ctx.close()
