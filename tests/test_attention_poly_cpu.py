"""The FMA-pipe exp2 of the attention kernel (`ex2_poly`, mebt_b200/csrc/attention.cu): a float32 NumPy replay of the
same arithmetic (round-to-nearest split through the 1.5 * 2^23 magic constant, degree-3 polynomial on [-0.5, 0.5], exponent
added as an integer) must stay within 1e-4 of 2^x over the whole range the softmax produces (x <= 8 after the lazy
stabiliser, clamped at -125).  bf16 rounding of P is 2^-9 = 2e-3, so the polynomial is invisible in the output."""
import re
from pathlib import Path

import numpy as np

SRC = (Path(__file__).resolve().parent.parent / "mebt_b200" / "csrc" / "attention.cu").read_text()


def _coefficients():
    body = SRC[SRC.index("float ex2_poly(float x)"):]
    body = body[:body.index("}")]
    c = [float(v) for v in re.findall(r"(\d\.\d+)f", body) if v not in ("125.", "12582912.")]
    c = [v for v in c if v < 2.0]
    assert len(c) == 4, c
    return [np.float32(v) for v in c]          # c3, c2, c1, c0 in order of appearance


def ex2_poly(x):
    c3, c2, c1, c0 = _coefficients()
    x = np.maximum(x.astype(np.float32), np.float32(-125.0))
    magic = np.float32(12582912.0)
    xf = (x + magic).astype(np.float32)
    f = (x - (xf - magic)).astype(np.float32)
    p = (c3 * f + c2).astype(np.float32)
    p = (p * f + c1).astype(np.float32)
    p = (p * f + c0).astype(np.float32)
    bits = p.view(np.int32) + (xf.view(np.int32) << 23)
    return bits.view(np.float32)


def test_polynomial_exp2_accuracy():
    x = np.concatenate([np.linspace(-125, 8, 400001), np.arange(-125, 9, dtype=np.float64), np.arange(-125, 8) + 0.5,
                        -np.logspace(-8, 2, 2000)]).astype(np.float32)
    got = ex2_poly(x).astype(np.float64)
    ref = np.exp2(x.astype(np.float64))
    rel = np.abs(got / ref - 1.0)
    assert rel.max() < 1e-4, rel.max()
    assert np.all(np.diff(ex2_poly(np.linspace(-20, 8, 100001).astype(np.float32))) >= -1e-7 * 256)   # monotone up to rounding
    below = ex2_poly(np.array([-1e4, -200.0, -126.0], dtype=np.float32))                              # clamped, tiny, finite
    assert np.all(np.isfinite(below)) and np.all(below < 1e-37)
