"""GPU tier: the dual launch (conv_dual.cuh) -- a residual block's 1x1 downsample conv computed as a second tile class
of the block's first 3x3 conv (torchvision BasicBlock as traced by the reference's convert_fx, quantization/engines.py:118).
The default parity tests already run with it (tensors and accumulators of both convs against the oracle); these tests pin
the launch structure and compare against the engine with every conv as a launch of its own.  Through the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from ievm_testutil import cached_quantized  # noqa: E402
from oracle import int8_forward as O  # noqa: E402
from oracle import model_factory as mf  # noqa: E402


def test_downsample_convs_ride_their_blocks_first_conv():
    import ievm_b200
    eng = ievm_b200.B200QuantizedResNet.from_converted(cached_quantized(mf.PRUNED_WIDTHS), max_batch=8)
    names = [L.name for L in eng.net.layers]
    fused = {names[i]: names[eng.layer_launch(i)] for i in range(len(names)) if eng.layer_launch(i) != i}
    assert fused == {"maxpool": "conv1", "layer2.0.downsample.0": "layer2.0.conv1",
                     "layer3.0.downsample.0": "layer3.0.conv1", "layer4.0.downsample.0": "layer4.0.conv1"}
    assert eng.launches_per_forward == 18         # front end + 16 conv launches + head (21 before the dual launch)
    eng.set_option("dual", 0)
    assert eng.launches_per_forward == 21
    assert all(eng.layer_launch(i) == i for i in range(2, len(names)))
    eng.close()


@pytest.mark.parametrize("widths", [mf.PRUNED_WIDTHS, mf.DEFAULT_CFG_WIDTHS, mf.UNPRUNED_WIDTHS], ids=["w57", "w60", "w64"])
def test_int8_dual_launch_equals_separate_launches_and_the_oracle(widths):
    """Batch 256 (every CTA runs tiles of both classes, ragged last round) and small ragged batches, every width set
    (resident and streamed weights, single CTAs and CTA pairs); tensors and accumulators of both convs of a dual launch
    against the oracle at batch 5."""
    import ievm_b200
    gm = cached_quantized(widths)
    eng = ievm_b200.B200QuantizedResNet.from_converted(gm, max_batch=256)
    x = mf.synthetic_images(256, seed=41).cuda()
    y_dual = {n: eng(x[:n]).clone() for n in (1, 5, 77, 256)}
    eng.set_option("dual", 0)
    for n, y in y_dual.items():
        assert torch.equal(y, eng(x[:n])), f"batch {n}"
    eng.set_option("dual", 1)
    eng.set_option("keep_tensors", 1)
    net = O.extract_qnet(gm)
    xs = mf.synthetic_images(5, seed=42)
    y = eng(xs.cuda()).cpu().numpy()
    yo = O.forward(net, xs.numpy(), keep=True)
    assert np.array_equal(y, yo)
    pairs = 0
    for i, L in enumerate(eng.net.layers):
        if L.op == 0 and i >= 2 and (eng.layer_launch(i) != i or (i + 1 < len(eng.net.layers) and eng.layer_launch(i + 1) == i)):
            assert np.array_equal(eng.conv_accumulators(L.name, 5), net.trace[L.name + ":acc"]), f"accumulators of {L.name}"
            tid = L.out_tensor
            assert np.array_equal(eng.read_tensor(tid), net.trace[eng.net.tensor_names[tid]]), f"tensor of {L.name}"
            pairs += 1
    assert pairs == 6
    eng.close()


def test_fp16_dual_launch_is_bitwise_the_separate_launches():
    """FP16 student: same k-block order into the same fp32 accumulators, so the two launch structures agree bitwise."""
    import ievm_b200
    for widths in (mf.PRUNED_WIDTHS, mf.UNPRUNED_WIDTHS):
        m16 = mf.cast_fp16(mf.make_student(widths))
        eng = ievm_b200.B200HalfResNet.from_half_module(m16, max_batch=256)
        assert eng.launches_per_forward == 18
        x = mf.synthetic_images(256, seed=43).half().cuda()
        y_dual = {n: eng(x[:n]).clone() for n in (3, 64, 256)}
        eng.set_option("dual", 0)
        assert eng.launches_per_forward == 21
        for n, y in y_dual.items():
            assert torch.equal(y, eng(x[:n])), f"batch {n}"
        eng.close()


def test_resnet50_has_no_dual_launches():
    """Bottleneck: conv1 is 1x1 stride 1 and the downsample samples the block input with the block's stride -- not a tap
    of any conv of the block, so every conv stays a launch of its own."""
    import ievm_b200
    t16 = mf.cast_fp16(mf.make_teacher())
    eng = ievm_b200.B200HalfResNet.from_half_module(t16, max_batch=2)
    assert all(eng.layer_launch(i) == i for i in range(2, len(eng.net.layers)))
    eng.close()
