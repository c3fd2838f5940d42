import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = ievm_b200.B200QuantizedResNet.from_converted(mf.static_quantize_fbgemm(mf.make_student()), max_batch=n)
x = mf.synthetic_images(n).cuda()
eng.set_option("use_graph", 1)
for _ in range(5):
    eng(x)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        eng(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"batch {n}: step {1e3 * e0.elapsed_time(e1) / 30:.1f} us")
