"""BASELINE.json config 3: pruned INT8 ResNet-18, batch sweep 1..4096 on one GPU (device-resident input, CUDA-graph
replay, CUDA events; the median of 5 blocks of `reps` forwards).  Writes one JSON object.
usage: batch_sweep.py out.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

out = sys.argv[1] if len(sys.argv) > 1 else "batch_sweep.json"
gm = mf.static_quantize_fbgemm(mf.make_student())
torch.backends.quantized.engine = "fbgemm"
rows = []
x_all = mf.synthetic_images(4096, seed=7)
for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=n)
    eng.set_option("use_graph", 1)
    x = x_all[:n].cuda()
    for _ in range(5):
        y = eng(x)
    torch.cuda.synchronize()
    reps = max(5, min(200, 4096 // n))
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            y = eng(x)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / reps)
    ms = sorted(times)[2]
    rows.append({"batch": n, "ms_per_forward": round(ms, 4), "images_per_s": round(n / ms * 1e3), "launches": eng.launches_per_forward})
    if n <= 64:      # spot parity against the reference's CPU module on the small batches
        with torch.no_grad():
            assert torch.equal(y.cpu(), gm(x_all[:n])), f"batch {n}: logits differ from CPU fbgemm"
    eng.close()
    print(rows[-1], flush=True)
json.dump({"workload": "int8_r18_pruned", "gpu": torch.cuda.get_device_name(0), "how": __doc__.split("\n")[0], "rows": rows},
          open(out, "w"), indent=1)
