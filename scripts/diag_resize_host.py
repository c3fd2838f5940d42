import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ievm_b200
from ievm_b200 import synthetic as mf
gm = mf.static_quantize_fbgemm(mf.make_student())
eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=8)
rng = np.random.default_rng(200)
imgs = rng.integers(0, 256, (3, 200, 200, 3), dtype=np.uint8)
u8 = torch.from_numpy(imgs)
d = eng.forward_u8(u8.cuda()).cpu()
h1 = eng.forward_u8(u8)
h2 = eng.forward_u8(u8)
hp = eng.forward_u8(u8.pin_memory())
d2 = eng.forward_u8(u8.cuda()).cpu()
print("dev==dev2", torch.equal(d, d2), "host1==dev", torch.equal(h1, d), "host2==dev", torch.equal(h2, d), "pinned==dev", torch.equal(hp, d), "h1==h2", torch.equal(h1, h2))
one = eng.forward_u8(u8[:1])
print("n=1 host==dev", torch.equal(one, d[:1]))
print((h1 - d).abs().max().item(), (h2 - d).abs().max().item())
