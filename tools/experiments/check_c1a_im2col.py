"""Desk check (CPU, numpy/torch) of the index arithmetic of the experimental tensor-core conv1a producer in tc_conv3x3.cuh:
im2col row r <-> halo pixel (hy, hx), tap k9 <-> (ky, kx) = (k9 / 3, k9 % 3) against the [64][9] weight layout, zero padding of the image,
and the zero mask for halo pixels outside the image (conv1b's padding).  Compares the emulated producer output for one 16x16 tile with
relu(conv3x3(img) + b) computed by torch on the whole image.  Not a substitute for the GPU run (tools/probe_k16.cu + the detector tests)."""
import numpy as np
import torch
import torch.nn.functional as F

rng = np.random.RandomState(3)
H = W = 64
img = rng.rand(H, W).astype(np.float16)
w1a = (rng.randn(64, 9) * 0.3).astype(np.float16)          # [out channel][ky*3+kx]  (OIHW with I = 1 flattened)
b1a = rng.randn(64).astype(np.float32)
ref = F.relu(F.conv2d(torch.from_numpy(img.astype(np.float32))[None, None], torch.from_numpy(w1a.astype(np.float32)).view(64, 1, 3, 3),
                      torch.from_numpy(b1a), padding=1))[0].numpy()          # [64][H][W]

HW, TH = 18, 16
for (tx, ty) in ((0, 0), (1, 2), (3, 3), (3, 0)):
    x0, y0 = tx * 16, ty * 16
    A = np.zeros((384, 16), np.float32)
    for r in range(HW * (TH + 2)):
        hy, hx = r // HW, r % HW
        cy, cx = y0 - 1 + hy, x0 - 1 + hx
        for k9 in range(9):
            iy, ix = cy + k9 // 3 - 1, cx + k9 % 3 - 1
            if 0 <= iy < H and 0 <= ix < W:
                A[r, k9] = img[iy, ix]
    Wm = np.zeros((64, 16), np.float32)
    Wm[:, :9] = w1a
    D = A @ Wm.T                                             # what the three MMAs produce (fp16 operands, fp32 accumulate)
    for r in range(HW * (TH + 2)):
        hy, hx = r // HW, r % HW
        oy, ox = y0 - 1 + hy, x0 - 1 + hx
        out = np.maximum(D[r] + b1a, 0)
        if not (0 <= oy < H and 0 <= ox < W):
            out = np.zeros(64, np.float32)                   # conv1b's zero padding
            exp = np.zeros(64, np.float32)
        else:
            exp = ref[:, oy, ox]
        assert np.abs(out - exp).max() < 1e-4, (tx, ty, r)
print("conv1a im2col index arithmetic: OK")
