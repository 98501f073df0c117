"""Wire formats of keys and proofs: the ``[@@deriving yojson]`` encodings of
``pkey`` / ``vkey`` / ``proof`` in the reference, so that keys and proofs written by a real
zukelang build can be loaded here and compared byte for byte, and the other way round.

What the reference emits (through ppx_yojson_conv and ``Yojson.Safe.to_string``):

* records -> JSON objects, fields in declaration order (groth16.ml:24-43,110-114;
  pinocchio.ml:37-75,195-208);
* ``'a list`` -> arrays; tuples -> arrays, so ``Var.t = string * int`` is ``["x",3]`` (var.ml:4-6);
* ``'a Var.Map.t`` -> the array of its bindings ``[[var, a], ...]`` in increasing key order
  (var.ml:38-40,66-68);
* ``Fr.t`` -> its value as a DECIMAL STRING (curve.ml:139-140 over misc.ml:36-38);
* ``G1.t`` / ``G2.t`` -> a JSON string holding the 48 / 96 COMPRESSED bytes verbatim
  (curve.ml:199-201,208-210: ``yojson_of_bytes`` is ```String (Bytes.to_string b)``);
* ``GT.t`` -> a JSON string holding ``GT.to_bytes`` (curve.ml:217-219).

A byte string is not UTF-8 in general, so documents are handled as ``bytes``.  The writer escapes
the way Yojson does: ``"`` ``\\`` ``\\b`` ``\\f`` ``\\n`` ``\\r`` ``\\t`` by name, every other byte
below 0x20 and 0x7f as ``\\u00xx`` (lower-case hex), and copies all other bytes — including those
>= 0x80 — unchanged; no whitespace is emitted.  The reader accepts any JSON whitespace and every
escape (``\\uXXXX`` becomes the UTF-8 encoding of the code point, as in Yojson's lexer).

Reading compressed points needs a square root per point: ``decode`` gathers every point of a
document and decompresses them with ONE ``zk_g*_decompress`` call per group.

Caveats (no OCaml toolchain in the build image): the byte-level claims above follow
the published behaviour of ppx_yojson_conv / Yojson, not output captured from the reference; and a
``GT`` blob written here is the library's own 576-byte encoding, not blst's, so ``vkey.ab`` does not
interoperate — ``Groth16Wire.vkey_of_yojson`` can recompute it from ``pkey.a`` and ``pkey.b2``.
"""

from __future__ import annotations

from typing import Any, Dict, List, Tuple

from .curve import Bls12_381, GTElem, Point, R

# ---- JSON over bytes ---------------------------------------------------------------------------
_NAMED = {0x22: b'\\"', 0x5C: b"\\\\", 0x08: b"\\b", 0x0C: b"\\f", 0x0A: b"\\n", 0x0D: b"\\r", 0x09: b"\\t"}
_UNNAMED = {v[1]: k for k, v in _NAMED.items()}          # escape letter -> byte
_UNNAMED[0x2F] = 0x2F                                    # "\/"


def _write_string(s: bytes, out: bytearray) -> None:
    out.append(0x22)
    for c in s:
        if c in _NAMED:
            out += _NAMED[c]
        elif c < 0x20 or c == 0x7F:
            out += b"\\u00%02x" % c
        else:
            out.append(c)
    out.append(0x22)


def _write(j: Any, out: bytearray) -> None:
    if isinstance(j, (bytes, bytearray)):
        _write_string(bytes(j), out)
    elif isinstance(j, str):
        _write_string(j.encode("utf-8"), out)
    elif isinstance(j, bool):
        out += b"true" if j else b"false"
    elif isinstance(j, int):
        out += b"%d" % j
    elif j is None:
        out += b"null"
    elif isinstance(j, dict):
        out.append(0x7B)
        for i, (k, v) in enumerate(j.items()):
            if i:
                out.append(0x2C)
            _write(k, out)
            out.append(0x3A)
            _write(v, out)
        out.append(0x7D)
    elif isinstance(j, (list, tuple)):
        out.append(0x5B)
        for i, v in enumerate(j):
            if i:
                out.append(0x2C)
            _write(v, out)
        out.append(0x5D)
    else:
        raise TypeError("wire.dumps: unsupported %r" % type(j))


def dumps(j: Any) -> bytes:
    """Compact JSON, strings as raw bytes (``Yojson.Safe.to_string``)."""
    out = bytearray()
    _write(j, out)
    return bytes(out)


class _Reader:
    def __init__(self, b: bytes):
        self.b, self.i = b, 0

    def fail(self, what: str):
        raise ValueError("wire.loads: %s at byte %d" % (what, self.i))

    def ws(self) -> None:
        b, n = self.b, len(self.b)
        while self.i < n and b[self.i] in b" \t\r\n":
            self.i += 1

    def peek(self) -> int:
        self.ws()
        if self.i >= len(self.b):
            self.fail("unexpected end")
        return self.b[self.i]

    def expect(self, c: int) -> None:
        if self.peek() != c:
            self.fail("expected %r" % chr(c))
        self.i += 1

    def string(self) -> bytes:
        self.expect(0x22)
        b, out = self.b, bytearray()
        while True:
            if self.i >= len(b):
                self.fail("unterminated string")
            c = b[self.i]
            self.i += 1
            if c == 0x22:
                return bytes(out)
            if c != 0x5C:
                out.append(c)
                continue
            if self.i >= len(b):
                self.fail("unterminated escape")
            e = b[self.i]
            self.i += 1
            if e in _UNNAMED:
                out.append(_UNNAMED[e])
            elif e == 0x75:                                # \uXXXX (surrogate pairs joined)
                cp = self._hex4()
                if 0xD800 <= cp < 0xDC00 and b[self.i:self.i + 2] == b"\\u":
                    self.i += 2
                    lo = self._hex4()
                    cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00)
                out += chr(cp).encode("utf-8", "surrogatepass")
            else:
                self.fail("bad escape")

    def _hex4(self) -> int:
        h = self.b[self.i:self.i + 4]
        if len(h) != 4:
            self.fail("bad \\u escape")
        self.i += 4
        try:
            return int(h, 16)
        except ValueError:
            self.fail("bad \\u escape")

    def value(self) -> Any:
        c = self.peek()
        if c == 0x22:
            return self.string()
        if c == 0x7B:
            self.i += 1
            d: Dict[bytes, Any] = {}
            if self.peek() == 0x7D:
                self.i += 1
                return d
            while True:
                k = self.string()
                self.expect(0x3A)
                d[k] = self.value()
                if self.peek() == 0x2C:
                    self.i += 1
                    continue
                self.expect(0x7D)
                return d
        if c == 0x5B:
            self.i += 1
            a: List[Any] = []
            if self.peek() == 0x5D:
                self.i += 1
                return a
            while True:
                a.append(self.value())
                if self.peek() == 0x2C:
                    self.i += 1
                    continue
                self.expect(0x5D)
                return a
        for lit, v in ((b"true", True), (b"false", False), (b"null", None)):
            if self.b.startswith(lit, self.i):
                self.i += len(lit)
                return v
        j = self.i
        while j < len(self.b) and self.b[j] in b"+-0123456789":
            j += 1
        if j == self.i:
            self.fail("unexpected byte")
        tok, self.i = self.b[self.i:j], j
        if self.i < len(self.b) and self.b[self.i] in b".eE":
            self.fail("floats do not occur in the wire format")
        return int(tok)


def loads(b: bytes) -> Any:
    """JSON -> dict (bytes keys) / list / bytes / int / bool / None."""
    r = _Reader(bytes(b))
    v = r.value()
    r.ws()
    if r.i != len(r.b):
        r.fail("trailing bytes")
    return v


# ---- typed codecs ------------------------------------------------------------------------------
# A schema is "Fr" | "G1" | "G2" | "GT" | ("list", T) | ("map", T) | ("record", ((field, T), ...)).
def List_(t):
    return ("list", t)


def Map_(t):
    return ("map", t)


def Record_(*fields):
    return ("record", tuple(fields))


def _get(v, name):
    return v[name] if isinstance(v, dict) else getattr(v, name)


def encode(schema, v, C=Bls12_381):
    """value -> JSON structure (feed to ``dumps``)."""
    if schema == "Fr":
        return b"%d" % (v % R)                                        # Z.to_string
    if schema == "G1":
        return C.G1.to_compressed_bytes(v)
    if schema == "G2":
        return C.G2.to_compressed_bytes(v)
    if schema == "GT":
        return C.GT.to_bytes(v)
    kind = schema[0]
    if kind == "list":
        return [encode(schema[1], x, C) for x in v]
    if kind == "map":
        return [[[k[0], k[1]], encode(schema[1], v[k], C)] for k in sorted(v)]
    if kind == "record":
        return {name: encode(t, _get(v, name), C) for name, t in schema[1]}
    raise TypeError("wire.encode: bad schema %r" % (schema,))


class _Pending:
    def __init__(self):
        self.g1: List[Point] = []
        self.g2: List[Point] = []


def _decode(schema, j, pend: _Pending):
    if schema == "Fr":
        if not isinstance(j, bytes):
            raise ValueError("wire.decode: Fr must be a decimal string")
        x = int(j)
        if not 0 <= x < R:
            x %= R                                                     # Fr.of_z reduces
        return x
    if schema in ("G1", "G2"):
        if not isinstance(j, bytes):
            raise ValueError("wire.decode: point must be a byte string")
        p = Point(b"", bytes(j))
        (pend.g1 if schema == "G1" else pend.g2).append(p)
        return p
    if schema == "GT":
        if not isinstance(j, bytes):
            raise ValueError("wire.decode: GT must be a byte string")
        return GTElem(j)
    kind = schema[0]
    if kind == "list":
        return [_decode(schema[1], x, pend) for x in j]
    if kind == "map":
        out = {}
        for binding in j:
            (name, idx), val = binding
            out[(name.decode("utf-8"), int(idx))] = _decode(schema[1], val, pend)
        return out
    if kind == "record":
        return {name: _decode(t, j[name.encode()], pend) for name, t in schema[1]}
    raise TypeError("wire.decode: bad schema %r" % (schema,))


def decode(schema, j, C=Bls12_381):
    """JSON structure (from ``loads``) -> value; records come back as dicts.  Points are
    decompressed — and checked: encoding, curve, subgroup — in one device call per group."""
    pend = _Pending()
    try:
        v = _decode(schema, j, pend)
    except (KeyError, IndexError, TypeError, AttributeError) as e:        # wrong shape for the schema
        raise ValueError("wire.decode: document does not match the schema (%r)" % (e,)) from None
    for G, pts in ((C.G1, pend.g1), (C.G2, pend.g2)):
        for p, full in zip(pts, G.of_compressed_bytes_many([p._comp for p in pts])):
            p.raw = full.raw
    return v


# ---- the reference's records -------------------------------------------------------------------
class Groth16Wire:
    """groth16.ml:24-43,110-114."""
    PKEY = Record_(("a", "G1"), ("d1", "G1"), ("ti1", List_("G1")), ("ltd_mid", Map_("G1")), ("tiztd", List_("G1")),
                   ("b1", "G1"), ("b2", "G2"), ("d2", "G2"), ("ti2", List_("G2")))
    VKEY = Record_(("one1", "G1"), ("ltgm_io", Map_("G1")), ("one2", "G2"), ("gm", "G2"), ("d", "G2"), ("ab", "GT"))
    PROOF = Record_(("a", "G1"), ("b", "G2"), ("c", "G1"))

    @staticmethod
    def yojson_of_pkey(pk) -> bytes:
        return dumps(encode(Groth16Wire.PKEY, pk))

    @staticmethod
    def yojson_of_vkey(vk) -> bytes:
        return dumps(encode(Groth16Wire.VKEY, vk))

    @staticmethod
    def yojson_of_proof(pr) -> bytes:
        return dumps(encode(Groth16Wire.PROOF, pr))

    @staticmethod
    def pkey_of_yojson(b: bytes):
        from .groth16 import PKey
        return PKey(**decode(Groth16Wire.PKEY, loads(b)))

    @staticmethod
    def vkey_of_yojson(b: bytes, recompute_ab_from=None):
        """``recompute_ab_from = pkey`` replaces the stored ``ab`` by ``e(pkey.a, pkey.b2)`` in this
        library's GT encoding (needed for a vkey written by the OCaml build)."""
        from .groth16 import VKey
        j = loads(b)
        if recompute_ab_from is not None:
            j = dict(j)
            j[b"ab"] = Bls12_381.GT.zero.raw
        vk = VKey(**decode(Groth16Wire.VKEY, j))
        if recompute_ab_from is not None:
            vk.ab = Bls12_381.Pairing.pairing(recompute_ab_from.a, recompute_ab_from.b2)
        return vk

    @staticmethod
    def proof_of_yojson(b: bytes):
        from .groth16 import Proof
        return Proof(**decode(Groth16Wire.PROOF, loads(b)))


class PinocchioWire:
    """pinocchio.ml:37-75,195-208."""
    PKEY = Record_(("vv", Map_("G1")), ("ww", Map_("G2")), ("yy", Map_("G1")), ("vav", Map_("G1")), ("waw", Map_("G2")),
                   ("yay", Map_("G1")), ("si", List_("G1")), ("bvwy", Map_("G1")), ("si2", List_("G2")), ("vt", "G1"),
                   ("wt", "G2"), ("yt", "G1"), ("vavt", "G1"), ("wawt", "G2"), ("yayt", "G1"), ("vbt", "G1"),
                   ("wbt", "G1"), ("ybt", "G1"), ("v_all", Map_("G1")), ("w_all", Map_("G1")))
    VKEY = Record_(("one", "G1"), ("one2", "G2"), ("av", "G2"), ("aw", "G1"), ("ay", "G2"), ("gm2", "G2"),
                   ("bgm", "G1"), ("bgm2", "G2"), ("yt", "G2"), ("vv_io", Map_("G1")), ("ww_io", Map_("G2")),
                   ("yy_io", Map_("G1")))
    PROOF = Record_(("vv", "G1"), ("ww", "G2"), ("yy", "G1"), ("h", "G1"), ("vavv", "G1"), ("waww", "G2"),
                    ("yayy", "G1"), ("bvwy", "G1"))

    @staticmethod
    def yojson_of_pkey(pk) -> bytes:
        return dumps(encode(PinocchioWire.PKEY, pk))

    @staticmethod
    def yojson_of_vkey(vk) -> bytes:
        return dumps(encode(PinocchioWire.VKEY, vk))

    @staticmethod
    def yojson_of_proof(pr) -> bytes:
        return dumps(encode(PinocchioWire.PROOF, pr))

    @staticmethod
    def pkey_of_yojson(b: bytes):
        from .pinocchio import PKey
        return PKey(**decode(PinocchioWire.PKEY, loads(b)))

    @staticmethod
    def vkey_of_yojson(b: bytes) -> dict:
        return decode(PinocchioWire.VKEY, loads(b))

    @staticmethod
    def proof_of_yojson(b: bytes):
        from .pinocchio import Proof
        return Proof(**decode(PinocchioWire.PROOF, loads(b)))


def yojson_of_solution(sol: Dict[Tuple[str, int], int]) -> bytes:
    """``Fr.t Var.Map.t`` (the public inputs handed to ``verify``)."""
    return dumps(encode(Map_("Fr"), sol))


def solution_of_yojson(b: bytes) -> Dict[Tuple[str, int], int]:
    return decode(Map_("Fr"), loads(b))
