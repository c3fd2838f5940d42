"""BASELINE.json config 5: KD evaluation step -- ResNet-50 teacher + (unpruned) ResNet-18 student FP16 forwards and
the soft-target loss of knowledge_distillation/train.py:47-57 -- on one GPU or under torchrun (global batch split
across ranks, loss terms all-reduced for reporting only).  Checks the loss against torch on rank 0, then times
`steps` evaluation steps with CUDA events.
    python scripts/kd_eval_bench.py [--batch 512] [--steps 10]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/kd_eval_bench.py --batch 512"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512, help="global batch")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--out", default="")
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
n = args.batch // world
teacher16 = mf.cast_fp16(mf.make_teacher())
student16 = mf.cast_fp16(mf.make_student(mf.UNPRUNED_WIDTHS))
teacher = ievm_b200.B200HalfResNet.from_half_module(teacher16, device=local, max_batch=n)
student = ievm_b200.B200HalfResNet.from_half_module(student16, device=local, max_batch=n)
for e in (teacher, student):
    e.set_option("use_graph", 1)
x = mf.synthetic_images(n, seed=7 + rank).half().to(dev)
y = torch.randint(0, 6, (n,), generator=torch.Generator().manual_seed(3 + rank)).to(dev)


def step():
    t = teacher(x)
    s = student(x)
    return ievm_b200.kd_eval_loss(s, t, y, alpha=0.5, temperature=4.0), s, t     # kd_config.py:17-18


for _ in range(3):
    out4, s, t = step()
torch.cuda.synchronize(dev)
if rank == 0:      # the loss against torch on the engine's own logits (train.py:47-57)
    T = 4.0
    sf, tf = s.float(), t.float()
    ce = torch.nn.functional.cross_entropy(sf, y)
    kd = torch.nn.KLDivLoss(reduction="batchmean")(torch.log_softmax(sf / T, 1), torch.softmax(tf / T, 1)) * T * T
    assert abs(float(out4[1]) - float(ce)) < 1e-3 and abs(float(out4[2]) - float(kd)) < 1e-3, (out4, ce, kd)
if dist is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(dev)
e0.record()
for _ in range(args.steps):
    out4, s, t = step()
e1.record()
torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
terms = out4.double().clone() * n          # sums over the local shard, for a global mean
if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(terms)
if rank == 0:
    line = {"workload": "kd_eval_fp16 (ResNet-50 teacher + ResNet-18 student + soft-target loss)", "n_gpus": world,
            "global_batch": n * world, "ms_per_step": float(ms), "images_per_s": n * world / float(ms) * 1e3,
            "loss": float(terms[0]) / (n * world), "ce": float(terms[1]) / (n * world), "kd": float(terms[2]) / (n * world),
            "launches_per_step": teacher.launches_per_forward + student.launches_per_forward + 2}
    print(json.dumps(line))
    if args.out:
        json.dump(line, open(args.out, "w"), indent=1)
if dist is not None:
    dist.destroy_process_group()
