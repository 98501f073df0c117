"""CPU oracle — BLS12-381 ate pairing in Python big integers (TEST INFRASTRUCTURE ONLY).

Stands in for ``Bls12_381.Pairing.pairing`` as the reference calls it at
``src/groth16/groth16.ml:103,168`` and ``src/pinocchio/pinocchio.ml:269`` so
the oracle can replay the reference's own acceptance test
(``assert (Protocol.verify ...)``, ``src/lib/test/test.ml:96,178``).

GT values are only ever compared with each other inside ``verify``; they are
not part of any proof, so this pairing need not be byte-compatible with blst's
GT encoding — it only has to be a non-degenerate bilinear map, which
``self_check`` verifies.  (It is the ate pairing with |z| as loop count, i.e.
blst's pairing up to inversion.)

Fp12 is Fp[w]/(w^12 - 2 w^6 + 2); Fp2 embeds by u -> w^6 - 1.  The sextic twist
maps (x, y) in E'(Fp2) to (x / w^2, y / w^3) in E(Fp12).  Line values are scaled
by w^3 (an element of Fp4, killed by the final exponentiation), which removes
every Fp12 inversion from the Miller loop.
"""

from __future__ import annotations

from .bls12_381 import P, R, G1, G2, f2_add, f2_sub, f2_mul, f2_inv

ATE_LOOP = 0xD201000000010000  # |z|
FINAL_EXP = (P ** 12 - 1) // R

F12_ONE = (1,) + (0,) * 11


def f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                if bj:
                    t[i + j] += ai * bj
    # w^12 = 2 w^6 - 2
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return tuple(x % P for x in t[:12])


def f12_pow(a, e):
    out = F12_ONE
    base = a
    while e:
        if e & 1:
            out = f12_mul(out, base)
        base = f12_mul(base, base)
        e >>= 1
    return out


def _embed(c, shift):
    """Fp2 element c = a + b u  ->  ((a - b) + b w^6) * w^shift as a sparse dict."""
    return {shift: (c[0] - c[1]) % P, shift + 6: c[1] % P}


def _line(lam, x1, y1, xp, yp):
    """w^3-scaled line through the twisted point (x1, y1) with Fp2 slope lam,
    evaluated at P = (xp, yp) in E(Fp):

        l = lam * xp * w^2  -  yp * w^3  +  (y1 - lam * x1)
    """
    coeffs = [0] * 12
    for k, v in _embed(f2_mul(lam, (xp, 0)), 2).items():
        coeffs[k] = (coeffs[k] + v) % P
    coeffs[3] = (coeffs[3] - yp) % P
    for k, v in _embed(f2_sub(y1, f2_mul(lam, x1)), 0).items():
        coeffs[k] = (coeffs[k] + v) % P
    return tuple(coeffs)


def miller_loop(p1, q2):
    """p1 in G1 (affine, Fp), q2 in G2 (affine, Fp2); neither may be the identity."""
    xp, yp = p1
    rx, ry = q2
    f = F12_ONE
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        # tangent at R
        three_x2 = f2_mul((3, 0), f2_mul(rx, rx))
        lam = f2_mul(three_x2, f2_inv(f2_add(ry, ry)))
        f = f12_mul(f12_mul(f, f), _line(lam, rx, ry, xp, yp))
        nx = f2_sub(f2_sub(f2_mul(lam, lam), rx), rx)
        ny = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
        rx, ry = nx, ny
        if (ATE_LOOP >> i) & 1:
            qx, qy = q2
            lam = f2_mul(f2_sub(qy, ry), f2_inv(f2_sub(qx, rx)))
            f = f12_mul(f, _line(lam, rx, ry, xp, yp))
            nx = f2_sub(f2_sub(f2_mul(lam, lam), rx), qx)
            ny = f2_sub(f2_mul(lam, f2_sub(rx, nx)), ry)
            rx, ry = nx, ny
    return f


def pairing(p1, q2):
    """e(p1, q2) in GT (tuple of 12 Fp coefficients)."""
    if p1 is None or q2 is None:
        return F12_ONE
    return f12_pow(miller_loop(p1, q2), FINAL_EXP)


def multi_pairing(pairs):
    """prod e(p_i, q_i) with one final exponentiation."""
    f = F12_ONE
    for p1, q2 in pairs:
        if p1 is None or q2 is None:
            continue
        f = f12_mul(f, miller_loop(p1, q2))
    return f12_pow(f, FINAL_EXP)


# GT as the additive-notation group the reference's ``GT`` module exposes
# (curve.ml:212-220 wraps it with ExtendG so ``+`` is the GT product).
def gt_add(a, b):
    return f12_mul(a, b)


def self_check() -> None:
    a, b = 0x1234567, 0x89ABCDE
    e = pairing(G1.one, G2.one)
    assert e != F12_ONE
    assert f12_pow(e, R) == F12_ONE
    assert pairing(G1.mul(G1.one, a), G2.mul(G2.one, b)) == f12_pow(e, a * b % R)
    assert multi_pairing([(G1.mul(G1.one, a), G2.one), (G1.neg(G1.one), G2.mul(G2.one, a))]) == F12_ONE


if __name__ == "__main__":
    import time
    t = time.time()
    self_check()
    print("oracle/pairing.py self-check OK in %.1fs" % (time.time() - t))
