"""GPU diagnostic: layout of a halo-patch TMA box in shared memory."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ievm_b200 import _lib
lib = _lib.load()
rng = np.random.default_rng(0)
for (h, w, cp, rb, bw, bh, w0, h0) in [(56, 56, 64, 64, 58, 6, -1, -1), (28, 28, 128, 128, 30, 8, -1, 3), (56, 56, 64, 128, 58, 6, -1, 53)]:
    x = rng.integers(1, 255, size=(2, h, w, cp), dtype=np.uint8)
    xd = torch.from_numpy(x).cuda()
    out = torch.zeros(bw * bh * rb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.ievm_probe_patch(xd.data_ptr(), 2, h, w, cp, rb, 1, w0, h0, bw, bh, out.data_ptr()), "probe_patch")
    got = out.cpu().numpy().reshape(bh * bw, rb)
    exp = np.zeros((bh, bw, rb), np.uint8)
    for r in range(bh):
        for c in range(bw):
            iy, ix = h0 + r, w0 + c
            if 0 <= iy < h and 0 <= ix < w:
                exp[r, c, :min(cp, rb)] = x[1, iy, ix, :rb]
    exp = exp.reshape(bh * bw, rb)
    chunks = rb // 16
    raw = np.empty_like(exp).reshape(-1, chunks, 16)
    e3 = exp.reshape(-1, chunks, 16)
    for i in range(exp.shape[0]):
        ph = (i % 8) if rb == 128 else ((i // 2) % 4)
        for j in range(chunks):
            raw[i, j ^ ph] = e3[i, j]
    raw = raw.reshape(-1, rb)
    print(f"patch h={h} w={w} cp={cp} rb={rb} box=({bw},{bh}) start=({w0},{h0}): swizzled-linear match={np.array_equal(got, raw)}, "
          f"unswizzled match={np.array_equal(got, exp)}, rows-as-sets match={np.array_equal(np.sort(got,1), np.sort(exp,1))}")
