"""GPU diagnostic: per-tensor / per-accumulator mismatch report of the INT8 engine against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ievm_b200
from ievm_b200 import synthetic as mf
from oracle import int8_forward as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
gm = mf.static_quantize_fbgemm(mf.make_student(mf.PRUNED_WIDTHS))
net = O.extract_qnet(gm)
eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=max(n, 2))
eng.set_option("keep_tensors", 1)
x = mf.synthetic_images(n, seed=21)
try:
    y = eng(x.cuda()).cpu().numpy()
except Exception as e:
    print("FORWARD FAILED:", e); sys.exit(1)
yo = O.forward(net, x.numpy(), keep=True)
print("env HALO=%s RB128=%s logits equal: %s" % (os.environ.get("IEVM_HALO"), os.environ.get("IEVM_HALO_RB128"), np.array_equal(y, yo)))
for tid, name in sorted(eng.net.tensor_names.items()):
    if name not in net.trace: continue
    got, ref = eng.read_tensor(tid), net.trace[name]
    bad = got != ref
    msg = "ok" if not bad.any() else "BAD %.4f%% first %s got %d ref %d maxdiff %d" % (
        100 * bad.mean(), tuple(int(v) for v in np.argwhere(bad)[0]), got[tuple(np.argwhere(bad)[0])], ref[tuple(np.argwhere(bad)[0])],
        np.abs(got.astype(int) - ref.astype(int)).max())
    print("tensor %-24s %s" % (name, msg))
    if bad.any() and "--all" not in sys.argv:
        # where are the bad pixels?
        b = bad.any(axis=1)
        ys, xs = np.where(b[0])
        print("   image0 bad rows", sorted(set(ys.tolist()))[:20], "bad cols", sorted(set(xs.tolist()))[:20], "bad ch", sorted(set(np.where(bad.any(axis=(0,2,3)))[0].tolist()))[:12])
        L = [l for l in eng.net.layers if l.out_tensor == tid][0]
        if L.op == 0 and (name + ":acc" in net.trace or L.name + ":acc" in net.trace):
            acc = eng.conv_accumulators(L.name, n); ra = net.trace[L.name + ":acc"]
            ba = acc != ra
            print("   acc of %s: bad %.4f%%" % (L.name, 100 * ba.mean()), "first", tuple(int(v) for v in np.argwhere(ba)[0]) if ba.any() else None,
                  (int(acc[tuple(np.argwhere(ba)[0])]), int(ra[tuple(np.argwhere(ba)[0])])) if ba.any() else "")
        break
