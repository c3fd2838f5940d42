"""One pass over every kernel family at small batch, for compute-sanitizer (memcheck / racecheck / synccheck):
    IEVM_WAIT_LIMIT_MS=0 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py
INT8 product configuration (fused front end, band-mode halo convs, im2col + CTA-pair convs, head), the un-fused parity
configuration, decoded 8-bit input with the resize stage, the pipelined host path, FP16 student and ResNet-50 teacher,
the loss / counter kernels.  Results are checked against the oracle so that a sanitizer-clean run is also a correct one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ievm_b200
from ievm_b200 import synthetic as mf
from oracle import int8_forward as O

gm = mf.static_quantize_fbgemm(mf.make_student(mf.PRUNED_WIDTHS))
eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=6)
x = mf.synthetic_images(5, seed=3)
want = O.forward(O.extract_qnet(gm), x.numpy())
assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)
assert np.array_equal(eng(x).numpy(), want)                              # host buffers
assert np.array_equal(eng.submit(x.pin_memory()).result().numpy(), want)  # pipelined host path
eng.set_option("use_graph", 1)
assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)
eng.set_option("use_graph", 0)
u8 = torch.randint(0, 256, (3, 200, 200, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
y_u8 = eng.forward_u8(u8.cuda()).cpu()
assert torch.equal(y_u8, eng.forward_u8(u8))
eng.set_option("keep_tensors", 1)                                          # un-fused front end + parity hooks
assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)
acc = eng.conv_accumulators("layer1.0.conv1", 5)
eng.set_option("conv_impl", 1)
assert np.array_equal(eng(x.cuda()).cpu().numpy(), want)
eng.close()

m16 = mf.cast_fp16(mf.make_student(mf.PRUNED_WIDTHS))
e16 = ievm_b200.B200HalfResNet.from_half_module(m16, max_batch=4)
x16 = mf.synthetic_images(3, seed=4).half().cuda()
with torch.no_grad():
    r16 = m16.cuda()(x16).float()
y16 = e16(x16).float()
assert float(((y16 - r16).abs().amax(1) / r16.abs().amax(1).clamp_min(1.0)).max()) < 1e-2
e16.close()
t16 = mf.cast_fp16(mf.make_teacher())
et = ievm_b200.B200HalfResNet.from_half_module(t16, max_batch=2)
xt = mf.synthetic_images(2, seed=5).half().cuda()
with torch.no_grad():
    rt = t16.cuda()(xt).float()
yt = et(xt).float()
assert float(((yt - rt).abs().amax(1) / rt.abs().amax(1).clamp_min(1.0)).max()) < 1e-2
out4 = ievm_b200.kd_eval_loss(y16[:2], yt, torch.tensor([1, 2]).cuda())
assert torch.isfinite(out4).all()
et.close()
torch.cuda.synchronize()
print("sanitize_smoke ok")
