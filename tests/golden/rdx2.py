"""Minimal reader for R's RDX2/RDX3 XDR serialisation (`save()` -> .rda).

Test infrastructure only: used by `make_fixtures.py` (run once, in the build
container, where /root/reference is mounted) to turn the reference's bundled
`data/*.rda` objects into the small `.npz` fixture committed next to this file.
R is not installed anywhere this repo runs, so the format is decoded by hand
(big-endian XDR; SEXP type in the low byte of the flags word; attributes and
tags signalled by flag bits 9/10; symbols kept in a back-reference table).
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP = 254, 253, 242, 241
REFSXP, NAMESPACESXP, PACKAGESXP, PERSISTSXP = 255, 249, 250, 247
MISSINGARG_SXP, UNBOUNDVALUE_SXP, BASENAMESPACE_SXP = 251, 252, 248
ALTREP_SXP, ATTRLISTSXP, ATTRLANGSXP = 238, 239, 240
SYMSXP, LISTSXP, CLOSXP, ENVSXP, LANGSXP = 1, 2, 3, 4, 6
CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP, VECSXP, EXPRSXP, RAWSXP = 9, 10, 13, 14, 15, 16, 19, 20, 24
NA_INT = -2147483648


class RObj:
    """A decoded R vector: `.value` (numpy array / list) + `.attr` dict."""

    def __init__(self, value, attr=None):
        self.value = value
        self.attr = attr or {}

    def names(self):
        n = self.attr.get("names")
        return None if n is None else list(n.value)

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP,
                 UNBOUNDVALUE_SXP, BASENAMESPACE_SXP):
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            self.i32()
            n = self.i32()
            v = [self.item() for _ in range(n)]
            self.refs.append(v)
            return v
        if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
            # pairlist: iterate over the cdr chain instead of recursing
            out = []
            attr = None
            while True:
                if t in (ATTRLISTSXP, ATTRLANGSXP):
                    has_attr = True
                if has_attr:
                    attr = self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t == NILVALUE_SXP:
                    break
                if t not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                    raise ValueError(f"unexpected cdr type {t}")
            return out
        if t == CHARSXP:
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("utf-8", "replace")
        if t in (LGLSXP, INTSXP):
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif t == REALSXP:
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t == CPLXSXP:
            n = self.length()
            v = np.frombuffer(self.raw(16 * n), dtype=">c16").astype(np.complex128)
        elif t == RAWSXP:
            n = self.length()
            v = np.frombuffer(self.raw(n), dtype=np.uint8).copy()
        elif t == STRSXP:
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t in (VECSXP, EXPRSXP):
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            self.item()  # attributes
            cls = info[0][1] if info else None
            if cls == "compact_intseq":
                n, start, step = (int(x) for x in state.value)
                return RObj(np.arange(start, start + n * step, step, dtype=np.int32))
            if cls == "wrap_integer" or cls == "wrap_real" or cls == "wrap_string":
                return state[0][1] if isinstance(state, list) else state
            raise ValueError(f"unsupported ALTREP class {cls}")
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        attr = {}
        if has_attr:
            for tag, val in self.item():
                attr[tag] = val
        return RObj(v, attr)


def load_rda(path):
    """Return {object name: RObj} for an .rda file written by R's save()."""
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    magic = raw[:5]
    if magic not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"not an RDX2/3 file: {magic!r}")
    r = _Reader(raw)
    r.p = 5
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR format supported")
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    return {tag: val for tag, val in top}


def matrix(obj):
    """R matrix (column-major + dim attr) -> numpy 2-D array, rownames, colnames."""
    d = obj.attr["dim"].value
    a = np.asarray(obj.value).reshape((int(d[0]), int(d[1])), order="F")
    rn = cn = None
    dn = obj.attr.get("dimnames")
    if dn is not None:
        rn = None if dn.value[0] is None else list(dn.value[0].value)
        cn = None if dn.value[1] is None else list(dn.value[1].value)
    return a, rn, cn
