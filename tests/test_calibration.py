"""CPU tier for SURVEY 8(f)-4 (calibration on the GPU): the host side -- flattening the prepared graph, the observer
plan, and the replay of per-batch (min, max) pairs into the prepared module's observers -- against the reference's own
CPU calibration loop (quantization/main.py:236-239).  The device reductions are covered by the gpu tier
(tests/test_zz_gpu_calibration.py)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ievm_b200
from ievm_b200 import calibration
from ievm_b200.netdesc import OP_ADD_RELU, OP_CONV, OP_HEAD, OP_MAXPOOL, POINT_LOGITS, POINT_POOLED
from oracle import model_factory as mf


def interpret(net, x):
    """fp32 torch evaluation of an FP16 NetSpec (weights as stored): checks the wiring of a flattened graph."""
    t = {0: x}
    for L in net.layers:
        if L.op == OP_CONV:
            y = F.conv2d(t[L.in_tensor], torch.from_numpy(L.weight).float(), torch.from_numpy(L.bias), L.stride, L.pad)
            if L.res_tensor >= 0:
                y = y + t[L.res_tensor]
            t[L.out_tensor] = torch.relu(y) if L.relu else y
        elif L.op == OP_MAXPOOL:
            t[L.out_tensor] = F.max_pool2d(t[L.in_tensor], 3, 2, 1)
        elif L.op == OP_ADD_RELU:
            t[L.out_tensor] = torch.relu(t[L.in_tensor] + t[L.res_tensor])
        else:
            pooled = t[L.in_tensor].mean((2, 3))
            return F.linear(pooled, torch.from_numpy(L.weight).float(), torch.from_numpy(L.bias)), t, pooled
    raise AssertionError("no head")


def recorded_stats(prepared, plan, num_tensors, batches, with_hist=False):
    """What the device logs would hold if activations were the reference's own fp32 ones: aminmax (and, for histogram
    observers, torch.histc over the running range of the point's observer group) of every observer call of the
    reference's CPU calibration forward, batch by batch, filed under the plan's observation points."""
    probe = copy.deepcopy(prepared)
    calls = []
    seen = set()
    for name, _ in plan:
        obs = probe.get_submodule(name)
        if id(obs) not in seen:                         # shared instances: one hook, fires once per call
            seen.add(id(obs))
            obs.register_forward_pre_hook(lambda m, inp: calls.append(inp[0].detach().clone()))
    groups = calibration.point_groups(prepared, plan, num_tensors)
    stats = np.full((len(batches), num_tensors + 2, 2), np.nan, np.float32)
    hists = np.zeros((len(batches), num_tensors + 2, calibration.HIST_BINS), np.uint32)
    run = {}
    with torch.no_grad():
        for r, (images, _) in enumerate(batches):
            calls.clear()
            probe(images)
            assert len(calls) == len(plan)
            cols = [calibration.point_index(point, num_tensors) for _, point in plan]
            for c, x in zip(cols, calls):
                stats[r, c] = [float(v) for v in torch.aminmax(x)]
                lo, hi = run.get(int(groups[c]), (np.inf, -np.inf))
                run[int(groups[c])] = (min(lo, stats[r, c, 0]), max(hi, stats[r, c, 1]))
            if with_hist:                                # the device bins after the whole batch's (min, max) pass
                for c, x in zip(cols, calls):
                    lo, hi = run[int(groups[c])]
                    hists[r, c] = torch.histc(x.float(), calibration.HIST_BINS, min=float(lo), max=float(hi)).numpy().astype(np.uint32)
    return (stats, hists) if with_hist else stats


def test_prepared_graph_flattens_to_an_unfused_fp16_net():
    m = mf.make_student((16, 24, 32, 40))
    prepared = mf.prepare_minmax(m)
    net, plan = ievm_b200.from_prepared(prepared, in_hw=(64, 64))
    ops = [L.op for L in net.layers]
    assert ops.count(OP_CONV) == 20 and ops.count(OP_ADD_RELU) == 8 and ops.count(OP_MAXPOOL) == 1 and ops[-1] == OP_HEAD
    assert all(L.res_tensor < 0 for L in net.layers if L.op == OP_CONV)          # nothing fused into a conv
    # 33 observer calls per forward on 30 distinct modules: (conv1, maxpool) and (last relu, avgpool, flatten) share
    assert len(plan) == 33 and len({id(prepared.get_submodule(n)) for n, _ in plan}) == 30
    assert plan[0][1] == 0 and plan[-1][1] == POINT_LOGITS and [p for _, p in plan].count(POINT_POOLED) == 2
    assert prepared.get_submodule(plan[1][0]) is prepared.get_submodule(plan[2][0])
    tensor_points = [p for _, p in plan if isinstance(p, int)]
    assert tensor_points == sorted(set(tensor_points)) and len(tensor_points) == 30      # every tensor observed once
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        ref = m(x)
    got, _, _ = interpret(net, x)
    assert torch.allclose(got, ref, atol=2e-2), float((got - ref).abs().max())   # weights rounded to fp16


@pytest.mark.parametrize("observer", ["moving_average", "minmax"])
def test_replaying_minmax_pairs_equals_the_reference_calibration(observer):
    """Feeding [min, max] of every observed tensor, batch by batch in call order, leaves the prepared module in exactly
    the state the reference's CPU loop leaves it in: identical observer state, identical converted network."""
    from torch.ao.quantization import QConfig, QConfigMapping, quantize_fx
    from torch.ao.quantization.observer import MinMaxObserver, PerChannelMinMaxObserver
    qmap = None
    if observer == "minmax":
        qc = QConfig(activation=MinMaxObserver.with_args(dtype=torch.quint8, reduce_range=True),
                     weight=PerChannelMinMaxObserver.with_args(dtype=torch.qint8, qscheme=torch.per_channel_symmetric))
        qmap = QConfigMapping().set_global(qc)
    m = mf.make_student((16, 24, 32, 40))
    gen = torch.Generator().manual_seed(5)
    batches = [(torch.randn(3, 3, 64, 64, generator=gen) * (1 + i), torch.zeros(3, dtype=torch.long)) for i in range(3)]
    ref = mf.prepare_minmax(m, qmap)
    ours = copy.deepcopy(ref)
    with torch.no_grad():
        for images, _ in batches:                       # quantization/main.py:236-239
            ref(images)
    net, plan = ievm_b200.from_prepared(ours, in_hw=(64, 64))
    calibration.check_observers(ours, plan)
    num_tensors = 1 + max(L.out_tensor for L in net.layers)
    stats = recorded_stats(ours, plan, num_tensors, batches)
    ievm_b200.replay_observers(ours, plan, stats, num_tensors)
    for name, _ in plan:
        a, b = ref.get_submodule(name), ours.get_submodule(name)
        assert torch.equal(a.min_val, b.min_val) and torch.equal(a.max_val, b.max_val), name
    qa = ievm_b200.from_converted(quantize_fx.convert_fx(ref), in_hw=(64, 64))
    qb = ievm_b200.from_converted(quantize_fx.convert_fx(ours), in_hw=(64, 64))
    assert (qa.in_scale, qa.in_zp) == (qb.in_scale, qb.in_zp) and len(qa.layers) == len(qb.layers)
    for la, lb in zip(qa.layers, qb.layers):
        for f in la.__dataclass_fields__:
            va, vb = getattr(la, f), getattr(lb, f)
            assert np.array_equal(va, vb) if isinstance(va, np.ndarray) else va == vb, (la.name, f)


def test_interpreter_statistics_match_the_plan_points():
    """The observation points of the plan are the tensors of the flattened net: aminmax of the interpreter's tensors
    equals what the reference's observers saw (same fp32 math up to fp16 weight rounding)."""
    m = mf.make_student((16, 24, 32, 40))
    prepared = mf.prepare_minmax(m)
    net, plan = ievm_b200.from_prepared(prepared, in_hw=(64, 64))
    x = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(9))
    num_tensors = 1 + max(L.out_tensor for L in net.layers)
    stats = recorded_stats(prepared, plan, num_tensors, [(x, None)])[0]
    logits, tensors, pooled = interpret(net, x)
    for _, point in plan:
        t = pooled if point == POINT_POOLED else logits if point == POINT_LOGITS else tensors[point]
        got = np.array([float(t.min()), float(t.max())])
        want = stats[calibration.point_index(point, num_tensors)]
        assert np.allclose(got, want, rtol=2e-2, atol=2e-2), (point, got, want)


def test_histc_restatement_equals_torch_histc():
    """oracle/observers.py (the formula the CUDA histogram kernel implements) against torch.histc itself."""
    from oracle.observers import histc_counts
    g = torch.Generator().manual_seed(0)
    for trial in range(24):
        n = int(torch.randint(1000, 200000, (1,), generator=g))
        x = torch.randn(n, generator=g) * float(torch.rand(1, generator=g) * 5 + 0.01)
        if trial % 2:
            x = torch.relu(x)
        if trial % 3 == 0:
            x = x.half().float()
        lo = float(x.min() - (torch.rand(1, generator=g) if trial % 4 == 0 else 0))
        hi = float(x.max() + (torch.rand(1, generator=g) * 2 if trial % 5 == 0 else 0))
        ref = torch.histc(x, 2048, min=lo, max=hi).numpy().astype(np.int64)
        assert np.array_equal(histc_counts(x.numpy(), lo, hi), ref), trial
    x = torch.full((100,), 3.0)
    assert np.array_equal(histc_counts(x.numpy(), 3.0, 3.0, 8), torch.histc(x, 8, min=3.0, max=3.0).numpy().astype(np.int64))


def test_replaying_histograms_equals_the_reference_fbgemm_calibration():
    """The default fbgemm qconfig of QuantizationEngine.static_quantize (quantization/engines.py:103: HistogramObserver,
    reduce_range): replaying per-batch (min, max) and histc counts over the observers' running ranges leaves the prepared
    module in exactly the state the reference's CPU loop (engines.py:123-133) leaves it in -- identical histograms,
    identical converted network."""
    from torch.ao.quantization import get_default_qconfig_mapping, quantize_fx
    from torch.ao.quantization.observer import HistogramObserver
    m = mf.make_student((16, 24, 32, 40))
    gen = torch.Generator().manual_seed(5)
    scales = (1.0, 2.5, 0.5, 2.5)                       # widening, then inside the running range
    batches = [(torch.randn(3, 3, 64, 64, generator=gen) * sc, torch.zeros(3, dtype=torch.long)) for sc in scales]
    ref = mf.prepare_minmax(m, get_default_qconfig_mapping("fbgemm"))
    ours = copy.deepcopy(ref)
    with torch.no_grad():
        for images, _ in batches:
            ref(images)
    net, plan = ievm_b200.from_prepared(ours, in_hw=(64, 64))
    assert calibration.observer_mode(ours, plan) == 2
    num_tensors = 1 + max(L.out_tensor for L in net.layers)
    groups = calibration.point_groups(ours, plan, num_tensors)
    assert groups[2] == 1 and groups[num_tensors] == groups[num_tensors - 1] == num_tensors - 1      # the shared instances
    stats, hists = recorded_stats(ours, plan, num_tensors, batches, with_hist=True)
    with pytest.raises(ValueError, match="histogram log"):
        ievm_b200.replay_observers(copy.deepcopy(ours), plan, stats, num_tensors)
    ievm_b200.replay_observers(ours, plan, stats, num_tensors, hists)
    for name, _ in plan:
        a, b = ref.get_submodule(name), ours.get_submodule(name)
        assert isinstance(a, HistogramObserver)
        assert torch.equal(a.min_val, b.min_val) and torch.equal(a.max_val, b.max_val), name
        assert torch.equal(a.histogram, b.histogram), name
    qa = ievm_b200.from_converted(quantize_fx.convert_fx(ref), in_hw=(64, 64))
    qb = ievm_b200.from_converted(quantize_fx.convert_fx(ours), in_hw=(64, 64))
    for la, lb in zip(qa.layers, qb.layers):
        assert (la.out_scale, la.out_zp, la.add_scale, la.add_zp) == (lb.out_scale, lb.out_zp, lb.add_scale, lb.add_zp), la.name


def test_other_observer_classes_are_refused_not_approximated():
    from torch.ao.quantization import QConfig, QConfigMapping
    from torch.ao.quantization.observer import PerChannelMinMaxObserver, default_weight_observer
    qc = QConfig(activation=PerChannelMinMaxObserver.with_args(dtype=torch.quint8, ch_axis=1), weight=default_weight_observer)
    prepared = mf.prepare_minmax(mf.make_student((16, 24, 32, 40)), QConfigMapping().set_global(qc))
    _, plan = ievm_b200.from_prepared(prepared)
    with pytest.raises(NotImplementedError, match="PerChannelMinMaxObserver"):
        calibration.check_observers(prepared, plan)


def test_main_py_flavour_is_accepted_and_the_oracle_matches_fbgemm():
    """The qconfig of quantization/main.py:187-222 (full 0..255 activation range, moving-average min/max) produces
    qparams the INT8 engine's descriptor accepts, and the oracle restatement stays bit-exact for it."""
    from oracle import int8_forward as O
    gm = mf.static_quantize_minmax(mf.make_student(mf.PRUNED_WIDTHS))
    net = ievm_b200.from_converted(gm)
    assert len(net.layers) == 22 and all(L.in_zp == 0 for L in net.layers[1:] if L.op == OP_CONV)
    x = mf.synthetic_images(4, seed=11)
    with torch.no_grad():
        ref = gm(x).numpy()
    assert np.array_equal(O.forward(O.extract_qnet(gm), x.numpy()), ref)
    sd = ievm_b200.from_quantized_state_dict(gm.state_dict())
    assert [(L.out_scale, L.out_zp, L.add_scale, L.add_zp) for L in sd.layers] == \
           [(L.out_scale, L.out_zp, L.add_scale, L.add_zp) for L in net.layers]
