"""Would two half-batch chains on two streams hide the launch boundaries?  Two engines (max_batch N/2 each), graph replay on
two torch streams, against one engine at batch N.  No kernel changes -- a probe for DESIGN section 7.
   python scripts/two_lane_probe.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ievm_b200
from ievm_b200 import synthetic as mf

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gm = mf.static_quantize_fbgemm(mf.make_student())
x = mf.synthetic_images(n).cuda()


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return 1e3 * best


one = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=n)
one.set_option("use_graph", 1)
y_ref = one(x).clone()
print(f"one engine, batch {n}: {timed(lambda: one(x)):.1f} us")

per = n // lanes
engs = [ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=per) for _ in range(lanes)]
streams = [torch.cuda.Stream() for _ in range(lanes)]
parts = [x[i * per:(i + 1) * per].contiguous() for i in range(lanes)]
for e in engs:
    e.set_option("use_graph", 1)
main = torch.cuda.current_stream()
outs = [None] * lanes


def both():
    ev = torch.cuda.Event()
    ev.record(main)
    for i in range(lanes):
        streams[i].wait_event(ev)
        with torch.cuda.stream(streams[i]):
            outs[i] = engs[i](parts[i])
        done = torch.cuda.Event()
        done.record(streams[i])
        main.wait_event(done)


t = timed(both)
torch.cuda.synchronize()
ok = torch.equal(torch.cat(outs), y_ref)
print(f"{lanes} engines x batch {per} on {lanes} streams: {t:.1f} us per {n} images   logits identical: {ok}")
