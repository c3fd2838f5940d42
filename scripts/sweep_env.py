"""Step time of the INT8 forward at batch 256 under one environment setting (graph replay, CUDA events), plus the
host-buffer e2e rate.  usage: VAR=val python scripts/sweep_env.py tag"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

tag = sys.argv[1] if len(sys.argv) > 1 else ""
n = 256
eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
xh = mf.synthetic_images(n).pin_memory()
x = xh.cuda()
eng.set_option("use_graph", 1)
for _ in range(5):
    y = eng(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40):
    y = eng(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
for _ in range(2):
    eng(xh)
t0 = time.perf_counter()
for _ in range(8):
    eng(xh)
e2e = 8 * n / (time.perf_counter() - t0)
print(f"[{tag}] step {ms*1000:.0f} us {n/ms*1000:.0f} img/s   e2e {e2e:.0f} img/s")
