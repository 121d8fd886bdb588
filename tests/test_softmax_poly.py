"""CPU emulation (numpy float32, the same operation order) of `ex2_poly2` in csrc/attention.cu — the exponentials the
attention softmax computes on the FMA pipe instead of MUFU.EX2 (one pair in four): round-to-nearest split x = n + f with
the 1.5 * 2^23 magic number, a degree-3 polynomial of 2^f on [-0.5, 0.5], n added to the exponent bits.  DESIGN.md §3.3 / §4
state a relative error of 2e-4, below the bf16 rounding (2^-9) of P."""
import numpy as np

C3, C2, C1, C0 = np.float32(0.053027521818876266), np.float32(0.24221394956111908), np.float32(0.6935725808143616), \
    np.float32(0.9999590516090393)
MAGIC = np.float32(12582912.0)


def ex2_poly(x):
    x = np.maximum(x.astype(np.float32), np.float32(-125.0))
    xr = (x + MAGIC).astype(np.float32)
    t = (xr - MAGIC).astype(np.float32)
    f = (t * np.float32(-1.0) + x).astype(np.float32)                 # fma(t, -1, x): exact here
    p = (f * C3 + C2).astype(np.float32)
    p = (p * f + C1).astype(np.float32)
    p = (p * f + C0).astype(np.float32)
    bits = p.view(np.int32) + (xr.view(np.int32) << np.int32(23))     # wraps modulo 2^32 like the device's integer add
    return bits.astype(np.int32).view(np.float32)


def test_relative_error_is_below_the_bf16_rounding_of_p():
    x = np.linspace(-60.0, 8.0, 400001, dtype=np.float64)             # softmax arguments: <= 8 (lazy rescale threshold)
    with np.errstate(over="ignore"):
        got = ex2_poly(x.astype(np.float32)).astype(np.float64)
    want = np.exp2(x.astype(np.float32).astype(np.float64))
    rel = np.abs(got - want) / want
    print(f"ex2_poly: max relative error {rel.max():.3e}")
    assert rel.max() <= 2.5e-4 < 2.0 ** -9 / 4


def test_masked_and_tiny_arguments_vanish():
    with np.errstate(over="ignore", invalid="ignore"):
        tiny = ex2_poly(np.array([-np.inf, -1e30, -126.0, -125.0], dtype=np.float32))
    assert np.all(tiny >= 0) and np.all(tiny <= 3e-38)                # masked keys contribute nothing to a row sum
    assert np.all(np.isfinite(tiny))


def test_exact_at_integers_up_to_the_constant_term():
    x = np.arange(-100, 9, dtype=np.float32)
    got = ex2_poly(x).astype(np.float64)
    assert np.allclose(got / np.exp2(x.astype(np.float64)), float(C0), rtol=1e-7)
