"""What each launch costs INSIDE the timed graph step: the step time with that launch left out (A/B build -DIEVM_EXP_SKIP,
IEVM_LIB_PATH=ab/lib_skip.so).  Per-launch CUDA events serialise the launches and so overstate small kernels; this is the
marginal cost with graph replay and programmatic dependent launch as in bench.py's timed region.
usage: IEVM_LIB_PATH=ab/lib_skip.so python scripts/skip_costs.py [N] [i8|f16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dtype = sys.argv[2] if len(sys.argv) > 2 else "i8"


def step_ms(mask, steps=30):
    os.environ["IEVM_SKIP_MASK"] = hex(mask)
    if dtype == "i8":
        eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=n)
    else:
        eng = ievm_b200.B200HalfResNet.from_half_module(gm, max_batch=n)
    eng.set_option("use_graph", 1)
    for _ in range(5):
        eng(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng(x)
    e1.record()
    torch.cuda.synchronize()
    names = [L.name for L in eng.net.layers]
    launch = [eng.layer_launch(i) for i in range(len(names))]
    eng.close()
    return e0.elapsed_time(e1) / steps, names, launch


if dtype == "i8":
    gm = mf.static_quantize_fbgemm(mf.make_student())
    x = mf.synthetic_images(n).cuda()
else:
    gm = mf.cast_fp16(mf.make_student())
    x = mf.synthetic_images(n).half().cuda()
base, names, launch = step_ms(0)
print(f"batch {n} {dtype}: full step {1e3 * base:.1f} us")
total = 0.0
for i, nm in enumerate(names):
    if launch[i] != i:
        continue
    mates = [names[j] for j in range(len(names)) if launch[j] == i and j != i]
    mask = 1 << i
    for j in range(len(names)):
        if launch[j] == i:
            mask |= 1 << j
    t, _, _ = step_ms(mask)
    total += base - t
    print(f"  without {nm:24s}{('(+' + ','.join(mates) + ')') if mates else '':28s} {1e3 * t:7.1f} us   marginal {1e3 * (base - t):6.1f} us")
print(f"sum of marginals {1e3 * total:.1f} us of {1e3 * base:.1f}")
