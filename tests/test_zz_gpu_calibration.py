"""GPU tier (-m gpu) for SURVEY 8(f)-4: PTQ calibration with the forwards and the observer reductions on the B200,
and the qconfig flavour of the reference's stage-4 script (quantization/main.py:187-222) on the INT8 engine.

Bars: the device (min, max) pairs are bit-identical to torch.aminmax of the very tensors the engine holds (min / max are
exact operations); the stand-alone add_relu layer is bit-identical to relu(a + b) rounded once; calibration as a whole
runs in fp16 where the reference runs fp32, so observer state agrees within CALIB_REL_TOL and the INT8 network converted
from it agrees with the reference-calibrated one on top-1 for >= TOP1_AGREE of the images (both stated here, measured
values are printed).  The INT8 engine itself stays bit-exact for whatever qparams it is given.
"""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import int8_forward as O  # noqa: E402
from oracle import model_factory as mf  # noqa: E402

CALIB_REL_TOL = 1.5e-2   # |observer range ours - reference| / reference range (fp16 vs fp32 activations; measured <= 1.06e-2)
TOP1_AGREE = 0.99        # top-1 agreement between the two INT8 networks (different qparams, same weights; measured 100 %)


def _prepared():
    return mf.prepare_minmax(mf.make_student(mf.PRUNED_WIDTHS))


def test_main_py_qconfig_flavour_is_bit_exact_on_the_engine():
    """Full-range (0..255) activations and large zero points on both add operands: logits, every tensor, bit-identical to
    the oracle (which equals live fbgemm for this flavour: tests/test_calibration.py) -- this is the case that needs
    ATen's fused dequantisation inside quantized::add_relu."""
    import ievm_b200
    gm = mf.static_quantize_minmax(mf.make_student(mf.PRUNED_WIDTHS))
    eng = ievm_b200.B200QuantizedResNet.from_converted(gm, device=0, max_batch=16)
    x = mf.synthetic_images(16, seed=11)
    qn = O.extract_qnet(gm)
    ref = O.forward(qn, x.numpy(), keep=True)
    assert np.array_equal(eng(x.cuda()).cpu().numpy(), ref)
    with torch.no_grad():
        assert np.array_equal(ref, gm(x).numpy())                       # and the live fbgemm module agrees
    eng.set_option("keep_tensors", 1)
    assert np.array_equal(eng(x.cuda()).cpu().numpy(), ref)
    for name in ("add_relu_2", "add_relu_4", "add_relu_6", "add_relu_7", "layer4.1.conv1"):
        assert np.array_equal(eng.read_tensor_by_name(name), qn.trace[name]), name
    eng.close()


def test_device_observers_equal_torch_aminmax_and_histc_of_the_engine_tensors():
    import ievm_b200
    from ievm_b200 import calibration
    from ievm_b200.netdesc import OP_ADD_RELU, POINT_LOGITS, POINT_POOLED
    prepared = _prepared()
    net, plan = ievm_b200.from_prepared(prepared)
    T = 1 + max(L.out_tensor for L in net.layers)
    groups = calibration.point_groups(prepared, plan, T)
    assert groups[2] == 1 and groups[T] == T - 1                        # (conv1, maxpool), (last relu, avgpool) share
    eng = ievm_b200.CalibrationEngine(net, mode=2, groups=groups, device=0, max_batch=8)
    assert eng.num_points == eng.num_tensors + 2 == T + 2 and eng.capacity == 64
    run = {}
    for r, (n, seed, gain) in enumerate(((5, 3, 1.7), (8, 4, 0.6), (3, 6, 2.4))):   # ragged; inside the range; widening
        x32 = (mf.synthetic_images(n, seed=seed) * gain).cuda()
        logits = eng(x32.half())
        eng.observe(x32)
        torch.cuda.synchronize()
        stats, hists = eng.read_observations(), eng.read_histograms()
        assert stats.shape == (r + 1, eng.num_points, 2) and hists.shape == (r + 1, eng.num_points, 2048)
        rec, hist = stats[-1], hists[-1]
        lg = logits.float().cpu()
        tensors = {0: x32.cpu(), T + 1: lg}
        for tid in range(1, T):
            t = eng.read_tensor(tid)                                     # NCHW over the real channels, f16
            assert t.shape[0] == n
            tensors[tid] = torch.from_numpy(t.astype(np.float32))
        for c, t in tensors.items():                                     # (min, max): exact
            assert rec[c, 0] == float(t.min()) and rec[c, 1] == float(t.max()), (c, net.tensor_names.get(c))
        for c in range(eng.num_points):                                  # the running range of every observer group
            lo, hi = run.get(int(groups[c]), (np.inf, -np.inf))
            run[int(groups[c])] = (min(lo, float(rec[c, 0])), max(hi, float(rec[c, 1])))
        for c, t in tensors.items():                                     # histc over that range: exact counts
            lo, hi = run[int(groups[c])]
            want = torch.histc(t, 2048, min=lo, max=hi).numpy().astype(np.int64)
            assert np.array_equal(hist[c].astype(np.int64), want), (r, c, net.tensor_names.get(c))
        assert int(hist[T].sum()) == n * net.layers[-1].cin              # avgpool output: every value binned once
        for L in net.layers:                                             # the un-fused residual add: one rounding
            if L.op == OP_ADD_RELU:
                want = torch.relu(tensors[L.in_tensor] + tensors[L.res_tensor]).half().float()
                assert torch.equal(tensors[L.out_tensor], want), L.name
        pooled = tensors[net.layers[-1].in_tensor].mean(dim=(2, 3))
        assert np.allclose(rec[T], [float(pooled.min()), float(pooled.max())], rtol=2e-3, atol=1e-3)
    # without the f32 batch the f16 input is observed; the log restarts with the option
    eng.reset_observations()
    x16 = mf.synthetic_images(4, seed=5).half().cuda()
    eng(x16)
    eng.observe()
    rec = eng.read_observations()
    assert rec.shape[0] == 1 and rec[0, 0, 0] == float(x16.min()) and rec[0, 0, 1] == float(x16.max())
    # a full log is drained to the host transparently
    for _ in range(eng.capacity + 3):
        eng.observe()
    more = eng.read_observations()
    assert more.shape[0] == eng.capacity + 4 and np.array_equal(more[-1], more[0])
    # the un-fused calibration net computes the same function as the fused FP16 engine (one extra rounding per block)
    fused = ievm_b200.B200HalfResNet.from_half_module(mf.cast_fp16(mf.make_student(mf.PRUNED_WIDTHS)), device=0, max_batch=8)
    a, b = eng(x16).float().cpu().numpy(), fused(x16).float().cpu().numpy()
    assert (np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1.0)).max() < 1e-2
    fused.close()
    eng.close()
    assert calibration.point_index(POINT_POOLED, 31) == 31 and calibration.point_index(POINT_LOGITS, 31) == 32


def test_gpu_calibration_matches_the_reference_cpu_calibration():
    """calibrate(prepared, loader) in place of the loop at quantization/main.py:236-239, then convert_fx as the reference."""
    import ievm_b200
    from torch.ao.quantization import quantize_fx
    calib = mf.calibration_batches(n_batches=3, batch=8, seed=1)
    ref = _prepared()
    ours = copy.deepcopy(ref)
    with torch.no_grad():
        for images, _ in calib:                                          # the reference's CPU calibration
            ref(images)
    stats = ievm_b200.calibrate(ours, calib, device=0, return_stats=True)
    assert stats.shape[0] == 3 and not np.isnan(stats).any()
    _, plan = ievm_b200.from_prepared(ours)
    worst = 0.0
    for name, _ in plan:
        a, b = ref.get_submodule(name), ours.get_submodule(name)
        rng = float(a.max_val - a.min_val)
        worst = max(worst, abs(float(b.min_val - a.min_val)) / rng, abs(float(b.max_val - a.max_val)) / rng)
    x0 = ours.get_submodule(plan[0][0]), ref.get_submodule(plan[0][0])
    assert torch.equal(x0[0].min_val, x0[1].min_val) and torch.equal(x0[0].max_val, x0[1].max_val)   # f32 input: exact
    print(f"observer state: worst relative deviation {worst:.2e}")
    assert worst < CALIB_REL_TOL
    gm_ref, gm_ours = quantize_fx.convert_fx(ref), quantize_fx.convert_fx(ours)
    assert float(gm_ours.conv1_input_scale_0) == float(gm_ref.conv1_input_scale_0)
    assert int(gm_ours.conv1_input_zero_point_0) == int(gm_ref.conv1_input_zero_point_0)
    na, nb = ievm_b200.from_converted(gm_ref), ievm_b200.from_converted(gm_ours)
    for la, lb in zip(na.layers, nb.layers):
        assert np.array_equal(la.weight, lb.weight) and np.array_equal(la.w_scale, lb.w_scale)       # weights: same path
        assert abs(lb.out_scale / la.out_scale - 1) < CALIB_REL_TOL and abs(lb.out_zp - la.out_zp) <= 3, la.name
    e_ref = ievm_b200.B200QuantizedResNet(na, device=0, max_batch=128)
    e_ours = ievm_b200.B200QuantizedResNet(nb, device=0, max_batch=128)
    x = mf.synthetic_images(128, seed=21)
    y_ref, y_ours = e_ref(x.cuda()).cpu().numpy(), e_ours(x.cuda()).cpu().numpy()
    assert np.array_equal(y_ours[:8], O.forward(O.extract_qnet(gm_ours), x[:8].numpy()))   # bit-exact for ITS qparams
    agree = float((y_ref.argmax(1) == y_ours.argmax(1)).mean())
    print(f"top-1 agreement GPU-calibrated vs CPU-calibrated INT8 network: {agree:.4f}; "
          f"max |logit difference| {np.abs(y_ref - y_ours).max():.3f}")
    assert agree >= TOP1_AGREE
    e_ref.close()
    e_ours.close()


def test_gpu_histogram_calibration_matches_the_reference_fbgemm_calibration():
    """The default flavour of QuantizationEngine.static_quantize (quantization/engines.py:95-121: default fbgemm qconfig,
    HistogramObserver): calibrate() in place of _calibrate (engines.py:123-133), convert_fx as the reference."""
    import ievm_b200
    from torch.ao.quantization import get_default_qconfig_mapping, quantize_fx
    calib = mf.calibration_batches(n_batches=3, batch=8, seed=1)
    ref = mf.prepare_minmax(mf.make_student(mf.PRUNED_WIDTHS), get_default_qconfig_mapping("fbgemm"))
    ours = copy.deepcopy(ref)
    with torch.no_grad():
        for images, _ in calib:
            ref(images)
    ievm_b200.calibrate(ours, calib, device=0)
    _, plan = ievm_b200.from_prepared(ours)
    x0 = ours.get_submodule(plan[0][0]), ref.get_submodule(plan[0][0])       # the f32 input: exact state
    assert torch.equal(x0[0].min_val, x0[1].min_val) and torch.equal(x0[0].histogram, x0[1].histogram)
    for name, _ in plan:
        a, b = ref.get_submodule(name), ours.get_submodule(name)
        # every element binned (float32 histograms: sums above 2^24 and the up-scaling of _combine_histograms round)
        assert abs(float(a.histogram.sum()) / float(b.histogram.sum()) - 1) < 1e-5, name
    na = ievm_b200.from_converted(quantize_fx.convert_fx(ref))
    nb = ievm_b200.from_converted(quantize_fx.convert_fx(ours))
    assert (na.in_scale, na.in_zp) == (nb.in_scale, nb.in_zp)
    worst = 0.0
    for la, lb in zip(na.layers, nb.layers):
        assert np.array_equal(la.weight, lb.weight)
        worst = max(worst, abs(lb.out_scale / la.out_scale - 1), abs(lb.add_scale / la.add_scale - 1) if la.res_tensor >= 0 else 0)
        assert abs(lb.out_zp - la.out_zp) <= 3, la.name
    print(f"histogram flavour: worst relative scale deviation {worst:.2e}")
    assert worst < 1.5e-2
    e_ref = ievm_b200.B200QuantizedResNet(na, device=0, max_batch=128)
    e_ours = ievm_b200.B200QuantizedResNet(nb, device=0, max_batch=128)
    x = mf.synthetic_images(128, seed=21).cuda()
    y_ref, y_ours = e_ref(x).cpu().numpy(), e_ours(x).cpu().numpy()
    agree = float((y_ref.argmax(1) == y_ours.argmax(1)).mean())
    print(f"histogram flavour: top-1 agreement GPU-calibrated vs CPU-calibrated INT8 network {agree:.4f}")
    assert agree >= TOP1_AGREE
    e_ref.close()
    e_ours.close()
