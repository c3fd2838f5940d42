"""Graph-replay step time of the INT8 pruned ResNet-18 (device events around 30 replays), for A/B runs:
   python scripts/step_time.py [N]      with IEVM_OVERLAP / IEVM_DUAL / IEVM_LIB_PATH in the environment"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf
ns = [int(a) for a in sys.argv[1:]] or [256]
gm = mf.static_quantize_fbgemm(mf.make_student())
for n in ns:
    eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=n)
    x = mf.synthetic_images(n).cuda()
    eng.set_option("use_graph", 1)
    for _ in range(5):
        eng(x)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            eng(x)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 30)
    print(f"batch {n}: step {1e3 * best:.1f} us  (overlap={os.environ.get('IEVM_OVERLAP', '1')})")
    eng.close()
