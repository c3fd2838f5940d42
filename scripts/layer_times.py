"""Per-launch device times (us) of one forward at batch N, plus graph-replay step time.
usage: layer_times.py [N] [tag] [i8|f16]   (IEVM_LIB_PATH / IEVM_* env vars select the variant)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tag = sys.argv[2] if len(sys.argv) > 2 else ""
if len(sys.argv) > 3 and sys.argv[3] == "f16":
    eng = ievm_b200.B200HalfResNet.from_half_module(mf.cast_fp16(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).half().cuda()
else:
    eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).cuda()
for _ in range(3):
    y = eng(x)
eng.set_option("profile", 1)
for _ in range(10):
    y = eng(x)
prof = {nm: round(ms / max(c, 1) * 1000) for nm, ms, c in eng.profile_read() if c}
eng.set_option("profile", 0)
eng.set_option("use_graph", 1)
for _ in range(5):
    y = eng(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30):
    y = eng(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 30
print(f"[{tag}] step {ms*1000:.0f} us  {n/ms*1000:.0f} img/s  checksum {float(y.sum()):.4f}")
print(f"[{tag}]", {k.replace('layer', 'L').replace('.conv', 'c').replace('downsample.0', 'ds'): v for k, v in prof.items() if v > 4})
