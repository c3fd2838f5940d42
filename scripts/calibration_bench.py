"""SURVEY 8(f)-4 measurement: the PTQ calibration loop (quantization/main.py:236-239, 256 calibration images as at
main.py:157) on the box's host cores, as the reference runs it, next to ievm_b200.calibrate (forwards + observer
reductions on the B200, host replay of the (min, max) pairs).  Also reports how far the two calibrations are apart.
    python scripts/calibration_bench.py [--images 256] [--batch 32] [--out file.json]"""
import argparse, copy, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--out", default="")
args = ap.parse_args()
torch.set_num_threads(os.cpu_count())
calib = mf.calibration_batches(n_batches=args.images // args.batch, batch=args.batch, seed=1)
ref = mf.prepare_minmax(mf.make_student(mf.PRUNED_WIDTHS))
ours = copy.deepcopy(ref)
warm = copy.deepcopy(ref)

with torch.no_grad():
    warm(calib[0][0])                       # thread pool / allocator warm-up outside the timed loop
    t0 = time.perf_counter()
    for images, _ in calib:
        ref(images.to("cpu"))
    cpu_s = time.perf_counter() - t0

torch.cuda.set_device(0)
pinned = [(x.pin_memory(), y) for x, y in calib]
ievm_b200.calibrate(copy.deepcopy(ours), pinned[:1], device=0)       # CUDA context, module load
torch.cuda.synchronize()
t0 = time.perf_counter()
stats = ievm_b200.calibrate(ours, pinned, device=0, return_stats=True)   # engine build + H2D + forwards + observers + replay
gpu_s = time.perf_counter() - t0

# the device part alone: batches resident on the GPU, one engine, events around forwards + observer passes
net, plan = ievm_b200.from_prepared(ours)
from ievm_b200 import calibration
groups = calibration.point_groups(ours, plan, 1 + max(L.out_tensor for L in net.layers))
dev_batches = [x.cuda() for x, _ in calib]
halves = [x.half() for x in dev_batches]


def device_only(mode):
    eng = ievm_b200.CalibrationEngine(net, mode=mode, groups=groups, device=0, max_batch=args.batch)
    for x32, x16 in zip(dev_batches[:2], halves[:2]):
        eng(x16); eng.observe(x32)
    eng.reset_observations()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    for x32, x16 in zip(dev_batches, halves):
        out = eng(x16)
        eng.observe(x32)
    e[1].record()
    for x32, x16 in zip(dev_batches, halves):          # the forwards alone, for the share of the observer passes
        out = eng(x16)
    e[2].record()
    torch.cuda.synchronize()
    total, fwd = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    # bytes one observer pass reads: every observed tensor once (f32 input, f16 tensors over their channel pitch)
    nbytes = 0
    for x32 in dev_batches:
        n = x32.shape[0]
        nbytes += x32.numel() * 4 + n * (net.layers[-1].cin + net.num_classes) * 2
        for tid in range(1, eng.num_tensors):
            _, h, w, _, pitch, elem = eng.tensor_shape(tid)
            nbytes += n * h * w * pitch * elem
    eng.close()
    passes = 1 if mode == 1 else 2
    return {"ms": total, "forward_ms": fwd, "observer_ms": total - fwd, "images_per_s": args.images / total * 1e3,
            "observer_bytes": nbytes * passes, "observer_gbs": nbytes * passes / max(total - fwd, 1e-6) / 1e6,
            "what": "forwards + observer passes (%s), batches resident in HBM, CUDA events" %
                    ("min/max" if mode == 1 else "min/max + histograms")}


dev1 = device_only(1)
dev2 = device_only(2)
dev_ms = dev1["ms"]

worst = 0.0
for name, _ in plan:
    a, b = ref.get_submodule(name), ours.get_submodule(name)
    rng = float(a.max_val - a.min_val)
    worst = max(worst, abs(float(b.min_val - a.min_val)) / rng, abs(float(b.max_val - a.max_val)) / rng)
line = {"workload": "PTQ calibration, pruned ResNet-18 [57,115,230,460], main.py qconfig (moving-average min/max)",
        "images": args.images, "batch": args.batch,
        "cpu_reference": {"seconds": cpu_s, "images_per_s": args.images / cpu_s, "threads": torch.get_num_threads(),
                          "what": "prepared_model(images) per batch on the host (quantization/main.py:236-239)"},
        "b200_calibrate": {"seconds": gpu_s, "images_per_s": args.images / gpu_s,
                           "what": "ievm_b200.calibrate from pinned host batches: engine build, H2D, fp16 forwards (un-fused "
                                   "adds, one buffer per tensor), device min/max passes, host replay"},
        "b200_device_only": dev1, "b200_device_only_histograms": dev2,
        "observer_state_worst_relative_deviation": worst, "records": int(stats.shape[0]), "points": int(stats.shape[1])}
print(json.dumps(line))
if args.out:
    json.dump(line, open(args.out, "w"), indent=1)
