"""The byte-level JSON layer of zukelang_b200/wire.py (no device needed): Yojson's escaping rules,
raw non-UTF-8 strings, the shapes ppx_yojson_conv gives records / tuples / Var.Map bindings."""
import json

import pytest

from zukelang_b200 import wire


def test_string_escaping_follows_yojson():
    s = bytes(range(256))
    enc = wire.dumps(s)
    assert enc[0] == 0x22 and enc[-1] == 0x22
    body = enc[1:-1]
    assert body.startswith(b"\\u0000\\u0001\\u0002\\u0003\\u0004\\u0005\\u0006\\u0007\\b\\t\\n\\u000b\\f\\r\\u000e")
    assert b"\\u001f !\\\"#" in body and b"[\\\\]" in body and b"~\\u007f\x80\x81" in body
    assert body.endswith(bytes(range(0x80, 0x100)))                   # high bytes are copied verbatim
    assert wire.loads(enc) == s


def test_ascii_documents_agree_with_the_json_module():
    doc = {"a": [1, -2, [["x", 3], "12345678901234567890123"]], "b": {"c": "q\"\\/\n"}, "d": [], "e": {}}
    enc = wire.dumps(doc)
    assert enc == json.dumps(doc, separators=(",", ":")).encode()      # compact, declaration order
    back = wire.loads(enc)
    assert back == {b"a": [1, -2, [[b"x", 3], b"12345678901234567890123"]], b"b": {b"c": b"q\"\\/\n"}, b"d": [], b"e": {}}
    spaced = json.dumps(doc, indent=2).encode()
    assert wire.loads(spaced) == back
    assert wire.loads(b'"\\u00e9\\ud83d\\ude00\\/"') == "é😀/".encode("utf-8")


@pytest.mark.parametrize("bad", [b"", b"{", b'{"a" 1}', b'"abc', b'"\\x"', b"[1,]", b"1.5", b"[1] 2", b'"\\u12"'])
def test_malformed_documents_are_rejected(bad):
    with pytest.raises(ValueError):
        wire.loads(bad)


def test_shapes_of_maps_and_fr():
    sol = {("x", 2): 5, ("ONE", 0): 1, ("x", 1): wire.R + 3}
    enc = wire.yojson_of_solution(sol)
    # Var.Map bindings in key order, Var.t as a 2-array, Fr as a decimal string (reduced)
    assert enc == b'[[["ONE",0],"1"],[["x",1],"3"],[["x",2],"5"]]'
    assert wire.solution_of_yojson(enc) == {("ONE", 0): 1, ("x", 1): 3, ("x", 2): 5}


def test_compression_of_uncompressed_bytes_matches_the_oracle():
    """_Group.to_compressed_bytes on a point that did not come from a device call: flag bits only."""
    import random
    from oracle import bls12_381 as O
    from zukelang_b200.curve import Bls12_381 as C, Point
    rng = random.Random(3)
    for G, Gm, unc, comp in ((O.G1, C.G1, O.g1_to_uncompressed, O.g1_compress),
                             (O.G2, C.G2, O.g2_to_uncompressed, O.g2_compress)):
        pts = [None, G.one, G.neg(G.one)] + [G.mul(G.one, rng.randrange(1, O.R)) for _ in range(10)]
        pts += [G.neg(p) for p in pts[3:]]
        for p in pts:
            assert Gm.to_compressed_bytes(Point(unc(p))) == comp(p)


def test_qap_build_matches_the_reference_restatement():
    """Host mirror of QAP.build (shared barycentric basis) against the oracle's literal restatement
    of QAP.ml:18-94 (one Polynomial.interpolate per variable) — no device needed."""
    from oracle import zk as Z
    from zukelang_b200 import qap as Q
    from zukelang_b200.protocol import Circuit, Gate, gate_set
    for (circ, _wit), literal in ((Z.circuit_cubic(), True), (Z.circuit_pair_case(7), True),
                                  (Z.circuit_mulchain(9), True), (Z.circuit_pair_case(40), False)):
        oq = Z.qap_build(circ.gates, literal=literal)
        gates = [Gate.make(dict(g.lhs), dict(g.l), dict(g.r)) for g in reversed(circ.gates)]   # any input order
        q, rgs = Q.build(gates + gates[:1])                                                     # duplicates collapse
        assert [r for r, _ in rgs] == list(range(len(circ.gates)))
        assert [(g.lhs, g.l, g.r) for _, g in rgs] == [(g.lhs, g.l, g.r) for g in circ.gates]   # Gate.Set order
        assert q.target == list(oq.target) and q.n == len(circ.gates)
        for name in ("v", "w", "y"):
            got, exp = getattr(q, name), getattr(oq, name)
            assert set(got) == set(exp)
            for k in exp:
                assert got[k] == list(exp[k]), (name, k)
        c = Circuit.of_gates(gates, circ.inputs_public, circ.outputs, circ.mids)
        assert list(c.vars) == circ.vars() and c.ios() == circ.ios() and list(c.gates) == gate_set(gates)


@pytest.mark.parametrize("doc", [b'{"a":"x"}', b'[1,2]', b'{"a":1,"b":2,"c":3}', b'"str"'])
def test_documents_of_the_wrong_shape_are_rejected(doc):
    with pytest.raises(ValueError):
        wire.decode(wire.Groth16Wire.PROOF, wire.loads(doc))
    with pytest.raises(ValueError):
        wire.decode(wire.Map_("Fr"), wire.loads(b'[["x","1"]]'))


def test_bench_circuit_generators_match_the_oracle():
    """tools/bench_groth16.py builds its synthetic circuits without the oracle (product-side code);
    they must be the circuits the parity tests prove on (oracle/zk.py generators)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_groth16 as BG
    from oracle import zk as Z
    for name, ogen in (("mulchain", Z.circuit_mulchain), ("r1cs", Z.circuit_random_r1cs)):
        for n in (2, 7, 64):
            oc, owit = ogen(n)
            pc, pwit = BG.CIRCUITS[name](n)
            assert [(dict(g.lhs), dict(g.l), dict(g.r)) for g in oc.gates] == pc.gates
            assert list(oc.inputs_public) == list(pc.inputs_public) and list(oc.outputs) == list(pc.outputs)
            assert sorted(oc.mids) == sorted(pc.mids) and oc.vars() == pc.variables
            assert owit(12345) == pwit(12345)
            assert Z.circuit_check(oc, pwit(99))
