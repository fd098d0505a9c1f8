"""ctypes wrapper of oracle/libcrossvar_ref.so (the oracle's C restatement; test infrastructure).
Build with `make -C oracle` (done by __graft_entry__.build())."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "libcrossvar_ref.so"))
        L.gmsref_pred_one_cross.restype = C.c_int
        L.gmsref_pred_one_cross.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                                               C.c_void_p, C.c_void_p,
                                                                               C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                                               C.c_void_p, C.c_void_p, C.c_int]
        L.gmsref_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def max_threads():
    return int(lib().gmsref_max_threads())


class RefProblem:
    """Inputs of the reference algorithm held in the layout the C code wants."""

    def __init__(self, H, add, dom=None, R=None, blocks=None, m_c=None, cM=None, chrom=None):
        """recombFreqMat as a dense matrix `R`, as its diagonal `blocks`, or as the map (`cM`, `chrom`): entries
        1-2*genmap2recombfreq() evaluated on the fly."""
        self.H = np.ascontiguousarray(H, dtype=np.uint8)
        self.m = self.H.shape[1]
        self.add = np.asfortranarray(add, dtype=np.float64)
        self.dom = None if dom is None else np.asfortranarray(dom, dtype=np.float64)
        self.T = self.add.shape[1]
        self.R = None if R is None else np.asfortranarray(R, dtype=np.float64)
        self.blocks = None
        if blocks is not None:
            self.blocks = [np.asfortranarray(b, dtype=np.float64) for b in blocks]
            self.off = np.concatenate([[0], np.cumsum([b.shape[0] for b in self.blocks])]).astype(np.int64)
            self.bptr = (C.c_void_p * len(self.blocks))(*[b.ctypes.data for b in self.blocks])
            assert int(self.off[-1]) == self.m
        self.cM = self.chrom = None
        if cM is not None:
            self.cM = np.ascontiguousarray(cM, dtype=np.float64)
            self.chrom = np.ascontiguousarray(chrom, dtype=np.int32)
            assert self.cM.size == self.m and self.chrom.size == self.m
        assert (self.R is not None) + (self.blocks is not None) + (self.cM is not None) == 1

    def one_cross(self, sire, dam, model="AD", nthreads=0):
        ad = 1 if model == "AD" else 0
        npairs, ncomp = self.T * (self.T + 1) // 2, 2 if ad else 1
        out = np.zeros((npairs, ncomp))
        nseg = C.c_int32(0)
        rows = [self.H[2 * sire], self.H[2 * sire + 1], self.H[2 * dam], self.H[2 * dam + 1]]
        rc = lib().gmsref_pred_one_cross(
            self.m, *[r.ctypes.data for r in rows],
            None if self.R is None else self.R.ctypes.data,
            0 if self.blocks is None else len(self.blocks),
            None if self.blocks is None else self.off.ctypes.data,
            None if self.blocks is None else C.cast(self.bptr, C.c_void_p),
            None if self.cM is None else self.cM.ctypes.data, None if self.chrom is None else self.chrom.ctypes.data,
            ad, self.T, self.add.ctypes.data, None if self.dom is None else self.dom.ctypes.data,
            out.ctypes.data, C.byref(nseg), int(nthreads))
        if rc != 0:
            raise MemoryError("gmsref_pred_one_cross: allocation failed")
        return out, int(nseg.value)
