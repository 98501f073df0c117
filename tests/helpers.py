"""Shared helpers for the parity tests: oracle <-> C-ABI byte conversions."""
import ctypes

from oracle import bls12_381 as O

R, P = O.R, O.P


def scalars_bytes(ks):
    return b"".join(O.fr_to_bytes(k) for k in ks)


def g1_bytes(pts):
    return b"".join(O.g1_to_uncompressed(p) for p in pts)


def g2_bytes(pts):
    return b"".join(O.g2_to_uncompressed(p) for p in pts)


def out_buf(n):
    return (ctypes.c_uint8 * n)()


def expect_g1(pt):
    return O.g1_to_uncompressed(pt) + O.g1_compress(pt)


def expect_g2(pt):
    return O.g2_to_uncompressed(pt) + O.g2_compress(pt)


def oracle_msm(G, pts, ks):
    """curve.ml:91-103 — left fold of scalar_mul + add."""
    acc = None
    for p, k in zip(pts, ks):
        acc = G.add(G.mul(p, k), acc)
    return acc


def random_points(G, n, rng, small=True):
    """n points with known discrete logs (dlogs returned too)."""
    dl = [rng.randrange(1, 1 << 20) if small else rng.randrange(1, R) for _ in range(n)]
    return [G.mul(G.one, d) for d in dl], dl


# ---- oracle <-> host-mirror conversions ------------------------------------------------
def mirror_point(pt, g2=False):
    from zukelang_b200.curve import Point
    return Point(O.g2_to_uncompressed(pt) if g2 else O.g1_to_uncompressed(pt))


def mirror_qap(oqap):
    from zukelang_b200.qap import QAP
    return QAP(v=dict(oqap.v), w=dict(oqap.w), y=dict(oqap.y), target=list(oqap.target))


def mirror_circuit(ocirc):
    from zukelang_b200.protocol import Circuit
    return Circuit(inputs_public=list(ocirc.inputs_public), outputs=list(ocirc.outputs), mids=list(ocirc.mids),
                   vars=ocirc.vars())


def mirror_groth16_pkey(opk):
    from zukelang_b200.groth16 import PKey
    p1 = lambda p: mirror_point(p)
    p2 = lambda p: mirror_point(p, True)
    return PKey(a=p1(opk.a), d1=p1(opk.d1), ti1=[p1(p) for p in opk.ti1], ltd_mid={k: p1(v) for k, v in opk.ltd_mid.items()},
                tiztd=[p1(p) for p in opk.tiztd], b1=p1(opk.b1), b2=p2(opk.b2), d2=p2(opk.d2), ti2=[p2(p) for p in opk.ti2])


def mirror_pinocchio_pkey(opk):
    from zukelang_b200.pinocchio import PKey
    g2_fields = {"ww", "waw", "si2", "wt", "wawt"}
    kw = {}
    for name, val in opk.items():
        g2 = name in g2_fields
        if isinstance(val, dict):
            kw[name] = {k: mirror_point(v, g2) for k, v in val.items()}
        elif isinstance(val, list):
            kw[name] = [mirror_point(v, g2) for v in val]
        else:
            kw[name] = mirror_point(val, g2)
    return PKey(**kw)


def groth16_proof_compressed(proof):
    return O.g1_compress(proof[0]) + O.g2_compress(proof[1]) + O.g1_compress(proof[2])


def pinocchio_proof_compressed(pr):
    from oracle import zk as Z
    return b"".join((O.g2_compress if g == "G2" else O.g1_compress)(pr[f])
                    for f, g in zip(Z.PINOCCHIO_PROOF_FIELDS, Z.PINOCCHIO_PROOF_GROUPS))


def decode_groth16_proof(proof):
    """host-mirror Proof -> oracle points (for the verifier replay)."""
    return (O.g1_from_uncompressed(proof.a.raw), O.g2_from_uncompressed(proof.b.raw), O.g1_from_uncompressed(proof.c.raw))


def decode_pinocchio_proof(proof):
    from zukelang_b200.pinocchio import PROOF_FIELDS, PROOF_IS_G2
    return {f: (O.g2_from_uncompressed if g2 else O.g1_from_uncompressed)(getattr(proof, f).raw)
            for f, g2 in zip(PROOF_FIELDS, PROOF_IS_G2)}


# ---- GT: C-ABI bytes (tower coefficients) <-> oracle/pairing.py tuples ---------------------------
def gt_bytes_to_oracle(b):
    """576 B of include/zkb200.h (slot a_j of c_i = the Fp2 coefficient x + y u of w^(2j+i), stored
    x then y, 48 B big-endian) -> the oracle's coefficients of w^0..w^11 with u = w^6 - 1."""
    assert len(b) == 576
    c = [int.from_bytes(b[48 * k:48 * k + 48], "big") for k in range(12)]
    t = [0] * 12
    o = 0
    for i in range(2):
        for j in range(3):
            k = 2 * j + i
            x, y = c[o], c[o + 1]
            o += 2
            t[k] = (x - y) % P
            t[k + 6] = y
    return tuple(t)


def oracle_to_gt_bytes(t):
    out = b""
    for i in range(2):
        for j in range(3):
            k = 2 * j + i
            y = t[k + 6]
            out += ((t[k] + y) % P).to_bytes(48, "big") + y.to_bytes(48, "big")
    return out
