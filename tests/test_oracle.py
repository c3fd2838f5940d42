"""CPU tier: pin the oracle (oracle/int8_forward.py) against (1) the golden fixtures recorded from the
reference's own QuantizationEngine (oracle/gen_golden.py) and (2) the live torch fbgemm operators."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import int8_forward as O
from oracle import model_factory as mf
from ievm_testutil import cached_quantized

WIDTHS = {"w57": mf.PRUNED_WIDTHS, "w60": mf.DEFAULT_CFG_WIDTHS, "w64": mf.UNPRUNED_WIDTHS}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("tag", ["w57", "w60", "w64"])
def test_oracle_matches_reference_golden(tag, golden_dir):
    g = np.load(os.path.join(golden_dir, f"int8_{tag}.npz"))
    gm = cached_quantized(WIDTHS[tag])
    net = O.extract_qnet(gm)
    # the regenerated synthetic model is the one the fixture was recorded from
    assert np.float32(net.in_scale) == g["in_scale"] and net.in_zp == int(g["in_zp"])
    x = mf.synthetic_images(int(g["n_images"])).numpy()
    logits = O.forward(net, x, keep=True)
    assert np.array_equal(logits, g["logits"]), "oracle logits differ from the reference engine's"
    names = [str(s) for s in g["node_names"]]
    digests = [str(s) for s in g["node_sha256"]]
    checked = 0
    for name, d in zip(names, digests):
        if name in net.trace:
            assert _sha(net.trace[name]) == d, f"activation digest mismatch at {name}"
            checked += 1
    assert checked >= 20
    for name in ("layer1.0.conv2", "layer3.0.downsample.0", "layer4.1.conv1"):
        assert np.array_equal(net.trace[name][0, :8, :4, :4], g[f"slice/{name}"])


def test_oracle_matches_the_main_py_flavour_golden(golden_dir):
    """Fixture of the stage-4 script's qconfig (quantization/main.py:185-242; oracle/gen_golden.py::gen_int8_minmax):
    full-range activations, zero points up to 158 -- the flavour on which quantized::add_relu's fused dequantisation
    (oracle/int8_forward.py::_dequant_fma) is observable."""
    g = np.load(os.path.join(golden_dir, "int8_minmax_w57.npz"))
    gm = mf.static_quantize_minmax(mf.make_student(mf.PRUNED_WIDTHS))
    net = O.extract_qnet(gm)
    assert np.float32(net.in_scale) == g["in_scale"] and net.in_zp == int(g["in_zp"])
    assert np.float32(net.blocks[6].add_scale) == g["scale/add_relu_6"] and net.blocks[6].add_zp == int(g["zp/add_relu_6"])
    logits = O.forward(net, mf.synthetic_images(int(g["n_images"])).numpy(), keep=True)
    assert np.array_equal(logits, g["logits"])
    checked = 0
    for name, d in zip((str(s) for s in g["node_names"]), (str(s) for s in g["node_sha256"])):
        if name in net.trace:
            assert _sha(net.trace[name]) == d, f"activation digest mismatch at {name}"
            checked += 1
    assert checked >= 20


def test_oracle_matches_live_fbgemm_nodes():
    gm = cached_quantized(mf.PRUNED_WIDTHS)
    net = O.extract_qnet(gm)
    x = mf.synthetic_images(3, seed=11)
    acts = {}
    hooks = [mod.register_forward_hook(lambda m_, i, o, name=name: acts.__setitem__(name, o))
             for name, mod in gm.named_modules() if name and not list(mod.children())]
    torch.backends.quantized.engine = "fbgemm"
    with torch.no_grad():
        y = gm(x)
    for hk in hooks:
        hk.remove()
    yo = O.forward(net, x.numpy(), keep=True)
    assert np.array_equal(y.numpy(), yo)
    for name, o in acts.items():
        if name in net.trace and getattr(o, "is_quantized", False):
            assert np.array_equal(o.int_repr().numpy().reshape(net.trace[name].shape), net.trace[name]), name


def test_oracle_ops_known_answers():
    # quantize: ties round to even, clamp to [0, 255]
    x = np.array([0.25, 0.75, -1e9, 1e9, 0.5], np.float32)
    assert O.quantize_input(x, 0.5, 3).tolist() == [3, 5, 0, 255, 4]   # rne(0.5)=0, rne(1.5)=2, rne(1)=1
    # maxpool ignores padding
    a = np.arange(16, dtype=np.uint8).reshape(1, 1, 4, 4)
    assert O.maxpool3x3s2(a)[0, 0].tolist() == [[5, 7], [13, 15]]
    # avgpool: rne(sum/49)
    b = np.full((1, 1, 7, 7), 10, np.uint8)
    b[0, 0, 0, 0] = 35          # sum = 515 -> 10.51 -> 11
    assert O.avgpool(b).tolist() == [[11]]
    # add_relu clamps at zero point 0 and at 255
    q = O.add_relu(np.array([0, 200], np.uint8), 1.0, 100, np.array([0, 200], np.uint8), 1.0, 0, 1.0, 0)
    assert q.tolist() == [0, 255]
    # integer conv with zero-point padding
    xq = np.full((1, 1, 3, 3), 7, np.uint8)
    w = np.ones((1, 1, 3, 3), np.int8)
    assert O.conv_acc(xq, 7, w, 1, 1).sum() == 0
    assert O.conv_acc(xq, 0, w, 1, 1)[0, 0].tolist() == [[28, 42, 28], [42, 63, 42], [28, 42, 28]]


def test_fp16_golden_is_sane(golden_dir):
    g = np.load(os.path.join(golden_dir, "fp16_w57.npz"))
    rel = np.abs(g["logits_fp16"] - g["logits_fp32"]).max(axis=1) / np.maximum(np.abs(g["logits_fp32"]).max(axis=1), 1)
    assert rel.max() < 1e-2


def test_pil_resize_restatement_matches_pillow():
    """oracle/pil_resize.py against PIL.Image.resize and torchvision's T.Resize (the reference's transform,
    quantization/dataset.py:15) on up- and down-scaling shapes: bit-identical."""
    from PIL import Image
    from torchvision import transforms as T
    from oracle.pil_resize import resize_bilinear_u8
    rng = np.random.default_rng(0)
    for h, w in [(200, 200), (224, 224), (300, 260), (199, 257), (64, 500), (97, 31)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
        assert np.array_equal(resize_bilinear_u8(img, 224, 224), ref), (h, w)
    img = rng.integers(0, 256, (200, 200, 3), dtype=np.uint8)          # NEU-DET images are 200 x 200
    assert np.array_equal(resize_bilinear_u8(img, 224, 224), np.asarray(T.Resize((224, 224))(Image.fromarray(img))))
