"""Minimal driver for profiling: build the INT8 engine at batch N and run a few forwards."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
workload = sys.argv[3] if len(sys.argv) > 3 else "int8"
if workload == "int8":
    eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).cuda()
elif workload == "fp16":
    eng = ievm_b200.B200HalfResNet.from_half_module(mf.cast_fp16(mf.make_student()), max_batch=n)
    x = mf.synthetic_images(n).half().cuda()
else:
    eng = ievm_b200.B200HalfResNet.from_half_module(mf.cast_fp16(mf.make_teacher()), max_batch=n)
    x = mf.synthetic_images(n).half().cuda()
for _ in range(iters):
    y = eng(x)
torch.cuda.synchronize()
print("checksum", float(y.float().sum()))
