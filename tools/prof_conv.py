"""Run a few representative launches of the tcgen05 kernels through the operator C ABI (authoring aid for ncu captures)."""
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


def conv(b, h, w, cin, cout, pool, reps=3):
    x = torch.randn(b, h, w, cin, device="cuda").half()
    wt = pack(torch.randn(cout, cin, 3, 3, device="cuda") * 0.05, 32 if cin == 32 else (cin + 63) // 64 * 64)
    bias = torch.randn(cout, device="cuda")
    out = torch.zeros(b, h, w, cout, dtype=torch.float16, device="cuda")
    po = torch.zeros(b, h // 2, w // 2, cout, dtype=torch.float16, device="cuda") if pool else None
    for _ in range(reps):
        capi.check(lib.airfe_op_conv3x3(x.data_ptr(), cin, w, h, b, cin, wt.data_ptr(), bias.data_ptr(), cout, cin, 1, out.data_ptr(), cout,
                                        po.data_ptr() if pool else None, cout, None))
    torch.cuda.synchronize()


which = sys.argv[1:] or ["c64", "c128", "c256"]
if "c64" in which:
    conv(4, 512, 512, 64, 64, True)
if "c128" in which:
    conv(4, 256, 256, 128, 128, False)
if "c256" in which:
    conv(4, 128, 128, 256, 320, False)
