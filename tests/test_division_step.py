"""The division of the batched LDLᵀ (csrc/batch.cuh, batch_div): quotient from
the correctly rounded reciprocal y = RN(1/d) by q0 = RN(w·y) and two Markstein
steps q ← RN(q + RN(w − q·d)·y). It must equal the IEEE quotient RN(w/d) — what
div.rn.f64 on the device and the reference's x86-64 division return — for every
operand pair in the range the kernel uses it for. Checked here in exact rational
arithmetic (float(Fraction) rounds to nearest-even) on random and adversarial
significands."""
from fractions import Fraction

import numpy as np


def rn(x: Fraction) -> float:
    return float(x)


def fma(a: float, b: float, c: float) -> float:
    return rn(Fraction(a) * Fraction(b) + Fraction(c))


def batch_div(w: float, d: float) -> float:
    y = rn(1 / Fraction(d))
    q = rn(Fraction(w) * Fraction(y))
    q = fma(fma(-q, d, w), y, q)
    q = fma(fma(-q, d, w), y, q)
    return q


def test_division_step_is_correctly_rounded():
    rng = np.random.default_rng(0)
    ulp = 2.0 ** -52
    special = [1.0, 1.0 + ulp, 2.0 - ulp, 2.0 - 2 * ulp, 1.5, 1.5 + ulp, 1.5 - ulp,
               4.0 / 3.0, 5.0 / 3.0, 1.0 + 2.0 ** -26, 1.0 + 2.0 ** -27 + ulp,
               1.9999999403953552, 1.4142135623730951, 1.7320508075688772]
    pairs = [(a, b) for a in special for b in special]
    pairs += [(a * 2.0 ** e1, -b * 2.0 ** e2) for a in special[:6] for b in special[:6]
              for e1 in (-40, 33) for e2 in (-30, 17)]
    for _ in range(40000):
        w = (1.0 + rng.random()) * 2.0 ** int(rng.integers(-60, 60)) * rng.choice([-1.0, 1.0])
        d = (1.0 + rng.random()) * 2.0 ** int(rng.integers(-60, 60)) * rng.choice([-1.0, 1.0])
        pairs.append((w, d))
    # significands next to the rounding boundaries of the quotient
    for _ in range(20000):
        d = float(1.0 + rng.integers(0, 2 ** 52) * ulp)
        q = float(1.0 + rng.integers(0, 2 ** 52) * ulp)
        w = q * d   # rounded product: w/d lands within an ulp of q
        pairs.append((w, d))
        pairs.append((np.nextafter(w, 4.0), d))
        pairs.append((np.nextafter(w, 0.0), d))
    bad = 0
    for w, d in pairs:
        want = rn(Fraction(float(w)) / Fraction(float(d)))
        got = batch_div(float(w), float(d))
        bad += got != want
    assert bad == 0
