"""Time representative launches of the tcgen05 conv kernel through the operator C ABI (authoring aid: CUDA-event timing and
ncu captures).  usage: prof_conv.py [c64|c64np|c64none|c32|c6432|c128|c256 ...]"""
import sys
import torch
sys.path.insert(0, ".")
from airslam_b200 import capi

lib = capi.lib()


def pack(w, c_pad):
    o, i, kh, kw = w.shape
    out = torch.zeros(o, kh * kw, c_pad, dtype=torch.float16, device="cuda")
    out[:, :, :i] = w.permute(0, 2, 3, 1).reshape(o, kh * kw, i).half()
    return out.reshape(o, -1).contiguous()


def conv(tag, b, h, w, cin, cout, full=True, pool=False, reps=5):
    x = torch.randn(b, h, w, cin, device="cuda").half()
    wt = pack(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05, cin if cin % 32 == 0 else (cin + 63) // 64 * 64)
    bias = torch.randn(cout, device="cuda")
    out = torch.zeros(b, h, w, cout, dtype=torch.float16, device="cuda") if full else None
    po = torch.zeros(b, h // 2, w // 2, cout, dtype=torch.float16, device="cuda") if pool else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        ev[i].record()
        capi.check(lib.airfe_op_conv3x3(x.data_ptr(), cin, w, h, b, cin, wt.data_ptr(), bias.data_ptr(), cout, cin, 1, out.data_ptr() if full else None, cout,
                                        po.data_ptr() if pool else None, cout, None))
    ev[reps].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(1, reps))
    fl = 2.0 * b * h * w * cout * 9 * cin
    print("%-10s %dx%dx%d %d->%d full=%d pool=%d : %.3f ms  %.0f TFLOP/s" % (tag, b, h, w, cin, cout, full, pool, ts[0], fl / ts[0] / 1e9))


which = sys.argv[1:] or ["c64", "c128", "c256"]
B = 16
for t in which:
    if t == "c64": conv(t, B, 512, 512, 64, 64, True, True)
    if t == "c64np": conv(t, B, 512, 512, 64, 64, True, False)
    if t == "c64po": conv(t, B, 512, 512, 64, 64, False, True)
    if t == "c6432": conv(t, B, 512, 512, 64, 32, True, False)
    if t == "c32": conv(t, B, 512, 512, 32, 32, False, True)
    if t == "c128": conv(t, B, 256, 256, 128, 128, True, False)
    if t == "c128po": conv(t, B, 256, 256, 128, 128, False, True)
    if t == "c256": conv(t, B, 128, 128, 256, 320, True, False)
    if t == "c96": conv(t, B, 256, 256, 96, 128, True, False)
    if t == "c12864": conv(t, 32, 128, 128, 128, 64, True, False)
    if t == "c12864s": conv(t, 32, 64, 64, 128, 64, True, False)
