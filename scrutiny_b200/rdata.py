"""Minimal writer / reader of the R workspace files MatriSeq.saveData produces next to each ``.npy``
(reference MatriSeq.py:1062-1065: ``save(<name>, file='<name>.gzip', compress=TRUE)`` of the array
converted by rpy2's numpy2ri) -- so ``R_files/data_load_methods.R:3-19`` keeps working without R or
rpy2 on the simulation host.

Format (R serialization version 2, XDR, gzip-compressed): ``RDX2\\nX\\n``, three version ints, then a
pairlist holding one tagged value: an INTSXP / REALSXP vector in column-major order with a ``dim``
attribute, closed by the NILVALUE marker.  Checked byte for byte against files written by R itself
(tests/golden/r_*.gzip, shipped by the reference under R_files/).
"""
import gzip
import struct

import numpy as np

_NIL, _SYM, _LIST, _CHAR, _INT, _REAL = 254, 1, 2, 9, 13, 14
_HAS_ATTR, _HAS_TAG = 1 << 9, 1 << 10
_R_VERSION, _R_MIN_VERSION = 0x00030400, 0x00020300    # what the reference's own files carry (R 3.4.0)
_NA_INT = -2 ** 31


def _charsxp(text):
    raw = text.encode("ascii")
    return struct.pack(">ii", (64 << 12) | _CHAR, len(raw)) + raw     # gp bit 6: ASCII


def _symbol(text):
    return struct.pack(">i", _SYM) + _charsxp(text)


def _vector(arr):
    arr = np.asarray(arr)
    flat = np.asarray(arr).reshape(-1, order="F")
    if np.issubdtype(arr.dtype, np.integer) or arr.dtype == np.bool_:
        body = struct.pack(">ii", _INT | _HAS_ATTR, flat.size) + flat.astype(">i4").tobytes()
    else:
        body = struct.pack(">ii", _REAL | _HAS_ATTR, flat.size) + flat.astype(">f8").tobytes()
    dims = np.asarray(arr.shape if arr.ndim else (1,), dtype=">i4")
    attr = (struct.pack(">i", _LIST | _HAS_TAG) + _symbol("dim") +
            struct.pack(">ii", _INT, dims.size) + dims.tobytes() + struct.pack(">i", _NIL))
    return body + attr


def serialize(name, arr):
    return (b"RDX2\nX\n" + struct.pack(">iii", 2, _R_VERSION, _R_MIN_VERSION) +
            struct.pack(">i", _LIST | _HAS_TAG) + _symbol(name) + _vector(arr) + struct.pack(">i", _NIL))


def write_rdata(path, name, arr):
    with gzip.open(path, "wb") as fh:
        fh.write(serialize(name, arr))


class _Reader:
    def __init__(self, raw):
        self.raw, self.pos = raw, 0

    def int(self):
        (v,) = struct.unpack_from(">i", self.raw, self.pos)
        self.pos += 4
        return v

    def take(self, n):
        out = self.raw[self.pos:self.pos + n]
        self.pos += n
        return out

    def symbol(self):
        assert self.int() & 0xFF == _SYM
        assert self.int() & 0xFF == _CHAR
        return self.take(self.int()).decode("ascii")

    def value(self):
        flags = self.int()
        kind = flags & 0xFF
        n = self.int()
        if kind == _INT:
            data = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int64)
        elif kind == _REAL:
            data = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        else:
            raise ValueError("unsupported SEXP type %d" % kind)
        dims = None
        if flags & _HAS_ATTR:
            while True:
                f = self.int()
                if f & 0xFF == _NIL:
                    break
                assert f & 0xFF == _LIST and f & _HAS_TAG
                tag = self.symbol()
                val = self.value()
                if tag == "dim":
                    dims = tuple(int(v) for v in val.reshape(-1))
        return data.reshape(dims, order="F") if dims else data


def read_rdata(path):
    """-> (name, ndarray) of a single-object file written by R's save() or by write_rdata."""
    with gzip.open(path, "rb") as fh:
        raw = fh.read()
    if raw[:7] != b"RDX2\nX\n":
        raise ValueError("not an RDX2 XDR workspace file")
    rd = _Reader(raw)
    rd.pos = 7
    version, _, _ = rd.int(), rd.int(), rd.int()
    assert version == 2
    f = rd.int()
    assert f & 0xFF == _LIST and f & _HAS_TAG
    name = rd.symbol()
    value = rd.value()
    assert rd.int() & 0xFF == _NIL
    return name, value
