"""The scalar arithmetic of the exact-digit evaluation kernel (facedeform_b200/csrc/fd_eval_tcx.cu), restated in numpy with the
constants READ FROM THE KERNEL SOURCE: exp2_digit (2^t from a 16-entry table and a quartic, the exponent added in the high
word) must hold the 2^-34 the error budget of DESIGN.md section 2 assumes, and split_digit (integer leading digit + remainder)
must be an exact decomposition.  No GPU: this pins the constants and the bit tricks, the kernel itself is tested under -m gpu
(reference loop evaluated: SOP_FaceDeform.cpp:404-439)."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "facedeform_b200", "csrc", "fd_eval_tcx.cu")


def _source_function(name):
    text = open(SRC).read()
    start = text.index(name + "(")
    return text[start:text.index("\n}\n", start)]


def _constants():
    body = _source_function("double exp2_digit")
    M = float(re.search(r"const double M = ([0-9.eE+-]+);", body).group(1))
    c4, c3 = (float(x) for x in re.search(r"double p = fma\(r, ([0-9.eE+-]+), ([0-9.eE+-]+)\);", body).groups())
    c2 = float(re.search(r"p = fma\(r, p, (0\.24[0-9]+)\);", body).group(1))
    c1 = float(re.search(r"p = fma\(r, p, (0\.69[0-9]+)\);", body).group(1))
    cut = int(re.search(r"> (0x[0-9A-Fa-f]+)u \? 0 : hi", body).group(1), 16)
    return M, (c1, c2, c3, c4), cut


def _hi(x):
    return (x.view(np.uint64) >> np.uint64(32)).astype(np.uint32)


def _lo(x):
    return (x.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def exp2_digit(t):
    """the kernel's sequence; a*b+c stands in for fma (the difference is 2^-53 relative, the claim is 2^-34)"""
    M, (c1, c2, c3, c4), cut = _constants()
    tab = np.exp2(np.arange(16) / 16.0)
    kf = t + M
    k = _lo(kf).astype(np.int32)
    r = t - (kf - M)
    p = r * c4 + c3
    p = r * p + c2
    p = r * p + c1
    T = tab[k & 15]
    s = (T * r) * p + T
    hi = (_hi(s).astype(np.int64) + ((k & ~np.int32(15)).astype(np.int64) << 16)) & 0xFFFFFFFF
    hi = np.where(_hi(t) > np.uint32(cut), 0, hi).astype(np.uint64)
    return ((hi << np.uint64(32)) | _lo(s).astype(np.uint64)).view(np.float64), r, s


def test_constants_are_the_taylor_coefficients_of_exp2():
    M, (c1, c2, c3, c4), cut = _constants()
    ln2 = np.log(2.0)
    assert M == 1.5 * 2.0 ** 48  # ulp 1/16: adding it rounds t to a multiple of 1/16 and leaves 16 t in the low word
    for c, n, f in ((c1, 1, 1.0), (c2, 2, 2.0), (c3, 3, 6.0), (c4, 4, 24.0)):
        assert abs(c - ln2 ** n / f) <= 1e-15 * c
    assert cut == 0xC0690000  # the high word of -200.0: below it the result is forced to (almost) zero


def test_exp2_digit_holds_2_to_minus_34():
    rng = np.random.default_rng(5)
    t = np.concatenate([rng.uniform(-199.9, 34.0, 2_000_000), rng.uniform(-30.0, 0.0, 2_000_000),
                        np.array([0.0, -0.0, 1 / 32, -1 / 32, -1 / 16, 30.0, -199.9, 1e-300, -1e-300])])
    y, r, s = exp2_digit(t)
    assert np.abs(r).max() <= 1.0 / 32.0
    assert s.min() > 0.97 and s.max() < 2.05  # so the exponent add in the high word cannot carry into the sign
    rel = np.abs(y - np.exp2(t)) / np.exp2(t)
    assert rel.max() <= 2.0 ** -34, rel.max()


def test_exp2_digit_vanishes_far_below_the_cut():
    t = np.array([-200.5, -1022.0, -5000.0, -1e300, -np.inf])
    with np.errstate(invalid="ignore"):  # -inf - (-inf) inside the range reduction: the high-word test discards the result
        y, _, _ = exp2_digit(t)
    assert np.all(y >= 0.0) and y.max() < 1e-300  # what is left is a denormal at most: nothing for a digit of 2^-11 and up


def test_split_digit_is_an_exact_decomposition():
    """x (|x| <= 2^11) = integer digit + remainder, |remainder| <= 1/2, the digit built by integer arithmetic on the float
    1.5 * 2^23 + k exactly as the kernel does (split_digit)."""
    body = _source_function("void split_digit")
    assert "0x4B400000" in body and "12582912.0f" in body
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-2048.0, 2048.0, 1_000_000), np.array([0.0, 0.5, -0.5, 1.5, 2047.5, -2047.5, 2048.0])])
    magic = 1.5 * 2.0 ** 52
    xm = x + magic
    k = _lo(xm).astype(np.int32)
    rem = (x - (xm - magic)).astype(np.float32)
    hi = (np.uint32(0x4B400000) + k.astype(np.uint32)).view(np.float32) - np.float32(12582912.0)
    assert np.all(hi == np.rint(x).astype(np.float32))  # round half to even, like the FP64 add
    assert np.all(np.abs(rem) <= 0.5)
    assert np.abs(hi.astype(np.float64) + rem.astype(np.float64) - x).max() <= 2.0 ** -25  # the FP32 rounding of the remainder
    assert np.all(hi.astype(np.float16).astype(np.float32) == hi)  # FP16 holds the digit exactly
