"""CPU tier: the arithmetic identities the CUDA epilogues rely on (csrc/conv_tc.cuh), checked in IEEE float32 with numpy
against the oracle's definitions (oracle/int8_forward.py), on millions of random and adversarial operands.

The kernels do not evaluate the oracle's formulas literally; they use
  * magic-number rounding: bits(x + 1.5 * 2^23) - bits(1.5 * 2^23) == rint(x) for |x| < 2^22   (round_add<kFast>)
  * integer clamps AFTER the rounding instead of float clamps before it (one VIADDMNMX.RELU each)
  * the ReLU of quantized::add_relu as a clamp of the rounded value at the zero point
  * ATen's fused dequantisation fma(scale, q, fl(scale * -zp)) for both operands of the add
and every one of those has to reproduce the reference bit for bit.  float64 holds the products and sums below exactly
(8/24-bit integers against 24-bit floats), so one cast to float32 is the fused multiply-add."""
import numpy as np
import pytest

from oracle import int8_forward as O

MAGIC = np.float32(12582912.0)          # 1.5 * 2^23
MAGIC_BITS = 0x4B400000


def bits(x):
    return np.asarray(x, np.float32).view(np.int32).astype(np.int64)


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def viaddmin_relu(x, add, mx):
    """max(min(x + add, mx), 0) on integers (__viaddmin_s32_relu)."""
    return np.maximum(np.minimum(x + add, mx), 0)


def test_magic_number_rounding_is_round_half_even_below_2_pow_21():
    rng = np.random.default_rng(0)
    x = np.concatenate([
        (rng.random(2_000_000, dtype=np.float32) - np.float32(0.5)) * np.float32(2 ** 22),      # |x| < 2^21
        (rng.integers(-2 ** 20, 2 ** 20, 500_000).astype(np.float32) + np.float32(0.5)),         # exact ties
        (rng.random(500_000, dtype=np.float32) - np.float32(0.5)) * np.float32(600.0),           # the u8 range
        np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -1.5, 255.5, 254.5, 2 ** 21 - 0.5], np.float32)])
    fast = bits(x + MAGIC) - MAGIC_BITS
    assert np.array_equal(fast, np.rint(x).astype(np.int64))


@pytest.mark.parametrize("relu", [False, True])
def test_requant_fast_form_equals_the_oracle(relu):
    """epilogue16_i8<kFast>: max(round_add(v) , lo) then saturating pack == clamp(rne(v) + zp, lo, 255)."""
    rng = np.random.default_rng(1)
    n, c = 400_000, 64
    acc = rng.integers(-160_000, 160_000, (n, c)).astype(np.int32)
    acc[:100] = rng.integers(-2 ** 20, 2 ** 20, (100, c))                                    # far outside the u8 range
    x_s, out_s, zp = 0.0123, 0.0456, 61
    w_s = (rng.random(c, dtype=np.float32) * np.float32(0.004) + np.float32(0.0005)).astype(np.float32)
    bias = (rng.standard_normal(c) * 0.5).astype(np.float32)
    want = O.requant(acc, x_s, w_s, bias, out_s, zp, relu, ch_axis=1)
    atw = (np.float32(x_s) * w_s).astype(np.float32)
    bdiv, mult = (bias / atw).astype(np.float32), (atw / np.float32(out_s)).astype(np.float32)
    v = ((acc.astype(np.float32) + bdiv).astype(np.float32) * mult).astype(np.float32)
    assert np.abs(v).max() < 2 ** 21                                                        # the bound the host checks
    q = bits(v + MAGIC) + (zp - MAGIC_BITS)
    q = np.maximum(q, zp if relu else 0)
    got = np.clip(q, 0, 255).astype(np.uint8)                                               # cvt.pack.sat.u8.s32
    assert np.array_equal(got, want)


def test_front_end_fast_form_with_both_clamps_in_one_instruction():
    """frontend_v2.cuh, kFast: the stem's ReLU output has zero point 0, so the lower clamp is 0 and
    clamp(rne(v) + zp, 0, 255) == max(min(bits(v + M) + (zp - M_bits), 255), 0) -- ONE VIADDMNMX.RELU (the host enables
    the instantiation only when out_lo == 0; also checked here for a non-zero zero point of a non-ReLU output)."""
    rng = np.random.default_rng(3)
    n, c = 400_000, 64
    acc = rng.integers(-200_000, 200_000, (n, c)).astype(np.int32)
    acc[:100] = rng.integers(-2 ** 20, 2 ** 20, (100, c))
    x_s, out_s = 0.0187, 0.0311
    w_s = (rng.random(c, dtype=np.float32) * np.float32(0.004) + np.float32(0.0005)).astype(np.float32)
    bias = (rng.standard_normal(c) * 0.5).astype(np.float32)
    atw = (np.float32(x_s) * w_s).astype(np.float32)
    bdiv, mult = (bias / atw).astype(np.float32), (atw / np.float32(out_s)).astype(np.float32)
    v = ((acc.astype(np.float32) + bdiv).astype(np.float32) * mult).astype(np.float32)
    assert np.abs(v).max() < 2 ** 21
    for zp, relu in ((0, True), (37, False)):                # lower clamp 0 in both cases
        want = O.requant(acc, x_s, w_s, bias, out_s, zp, relu, ch_axis=1)
        got = viaddmin_relu(bits(v + MAGIC), zp - MAGIC_BITS, 255).astype(np.uint8)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("flavour", ["fbgemm_reduce_range", "main_py_full_range"])
def test_fused_add_relu_fast_form_equals_the_oracle(flavour):
    """epilogue16_i8_res<kFast>: requantise the conv, dequantise it and the residual with ATen's fma form, add, ReLU,
    requantise -- with integer clamps after magic rounding -- against oracle requant + add_relu."""
    rng = np.random.default_rng(2)
    n, c = 300_000, 64
    acc = rng.integers(-120_000, 120_000, (n, c)).astype(np.int32)
    res = rng.integers(0, 256, (n, c)).astype(np.uint8)
    if flavour == "fbgemm_reduce_range":
        x_s, out_s, out_zp, r_s, r_zp, add_s, add_zp = 0.0119, 0.0391, 67, 0.0208, 58, 0.0311, 0
    else:                                   # quantization/main.py:187-222: both zero points large (layer4.0's add)
        x_s, out_s, out_zp, r_s, r_zp, add_s, add_zp = 0.0061, 0.0197, 133, 0.0101, 117, 0.0159, 0
    w_s = (rng.random(c, dtype=np.float32) * np.float32(0.004) + np.float32(0.0005)).astype(np.float32)
    bias = (rng.standard_normal(c) * 0.5).astype(np.float32)
    a_q = O.requant(acc, x_s, w_s, bias, out_s, out_zp, False, ch_axis=1)                   # conv2: no ReLU of its own
    want = O.add_relu(a_q, out_s, out_zp, res, r_s, r_zp, add_s, add_zp)
    # ---- the kernel's form ----
    atw = (np.float32(x_s) * w_s).astype(np.float32)
    bdiv, mult = (bias / atw).astype(np.float32), (atw / np.float32(out_s)).astype(np.float32)
    t = ((acc.astype(np.float32) + bdiv).astype(np.float32) * mult).astype(np.float32)
    out_lo = 0
    c1_add, c1_max = -MAGIC_BITS - (out_lo - out_zp), 255 - out_lo
    tq = viaddmin_relu(bits(t + MAGIC), c1_add, c1_max)                                     # q2 - out_lo
    a_scale, r_scale = np.float32(out_s), np.float32(r_s)
    pa = np.float32(a_scale * np.float32(-out_zp))
    pb = np.float32(r_scale * np.float32(-r_zp))
    a = fma(a_scale, (tq.astype(np.float32) + np.float32(out_lo)).astype(np.float32), pa)
    rf = ((res.astype(np.int64) | MAGIC_BITS).astype(np.int32).view(np.float32) - MAGIC).astype(np.float32)   # PRMT + FADD
    assert np.array_equal(rf, res.astype(np.float32))
    rb = fma(r_scale, rf, pb)
    inv = np.float32(1.0) / np.float32(add_s)
    u = ((a + rb).astype(np.float32) * inv).astype(np.float32)
    q = viaddmin_relu(bits(u + MAGIC), -MAGIC_BITS, 255 - add_zp) + add_zp
    assert np.array_equal(q.astype(np.uint8), want)
