"""Quick wall-clock timing of the public detect call (authoring aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from airslam_b200 import capi
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
imgs = np.stack([synth.stereo_pair(752, 480, 100 + i)[0] for i in range(B)])
ctx = capi.Context(max_batch=B, enable_lightglue=0)
for net, name, kw in ((capi.NET_SUPERPOINT, "superpoint", {}), (capi.NET_PLNET, "plnet", dict(lines=True, junctions=True))):
    for _ in range(3):
        ctx.detect_batch(net, imgs, **kw)
    t = time.time()
    n = 10
    for _ in range(n):
        r = ctx.detect_batch(net, imgs, **kw)
    dt = (time.time() - t) / n
    print("%s: batch %d  %.2f ms/batch  %.3f ms/image  (%d kpts, %s lines)" % (name, B, dt * 1e3, dt * 1e3 / B, r[0][0].shape[1], None if r[0][1] is None else len(r[0][1])))
