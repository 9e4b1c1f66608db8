"""Time representative launches of the generic tcgen05 GEMM kernel through the operator C ABI (authoring aid).
usage: prof_gemm.py [qkv|proj|pw128_256|pw256_128|pw320_9|pw256_256 ...]   rows are laid out as (W=rows per slot, H=1, B=slots)."""
import sys
import torch
sys.path.insert(0, ".")
from airslam_b200 import capi

lib = capi.lib()


def gemm(tag, w, h, b, cin, cout, block_n, tile, reps=6, out_f32=False):
    x = torch.randn(b, h, w, cin, device="cuda").half()
    wt = (torch.randn(cout, cin, device="cuda") * 0.05).half().contiguous()
    bias = torch.randn(cout, device="cuda")
    out = torch.zeros(b, h, w, cout, dtype=torch.float32 if out_f32 else torch.float16, device="cuda")
    tw, th, tb = tile
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    for i in range(reps):
        ev[i].record()
        capi.check(lib.airfe_op_tc_gemm(x.data_ptr(), cin, w, h, b, cin, w * cin, h * w * cin,
                                        wt.data_ptr(), cin, cout, cin, 0, 0, 0,
                                        1, cin, block_n, bias.data_ptr(), 0, int(out_f32),
                                        out.data_ptr(), h * w * cout, w * cout, cout, cout, tw, th, tb, None))
    ev[reps].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(1, reps))
    fl = 2.0 * b * h * w * cout * cin
    by = b * h * w * (cin * 2 + cout * out.element_size())
    print("%-10s rows=%d %d->%d bn=%d : %.4f ms  %.0f TFLOP/s  %.2f TB/s" % (tag, b * h * w, cin, cout, block_n, ts[0], fl / ts[0] / 1e9, by / ts[0] / 1e9))


which = sys.argv[1:] or ["qkv", "proj", "pw128_256", "pw256_128", "pw320_9", "pw256_256"]
for t in which:
    if t == "qkv": gemm(t, 512, 1, 32, 256, 768, 256, (128, 1, 1))
    if t == "qkv128": gemm(t, 512, 1, 32, 256, 768, 128, (128, 1, 1))
    if t == "proj": gemm(t, 512, 1, 32, 256, 256, 256, (128, 1, 1))
    if t == "proj128": gemm(t, 512, 1, 32, 256, 256, 128, (128, 1, 1))
    if t == "pw128_256": gemm(t, 128, 128, 32, 128, 256, 256, (16, 8, 1))
    if t == "pw256_128": gemm(t, 128, 128, 32, 256, 128, 128, (16, 8, 1))
    if t == "pw320_9": gemm(t, 128, 128, 32, 320, 16, 16, (16, 8, 1))
    if t == "pw256_256": gemm(t, 64, 64, 32, 256, 256, 256, (16, 8, 1))
