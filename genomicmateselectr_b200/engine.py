"""CrossVarEngine: a thin object over one `gms_handle` (one GPU).

Holds the HBM-resident state that is shared by the many predCrossVars calls of a
cross-validation run: per-chromosome LD-decay tile images R and RoR and the
bit-packed phased haplotypes.  All arithmetic happens in libgms_b200.so."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import GmsError, MODEL


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def trait_pairs(T):
    """Pair order of the result slab: variances first, then combn(traits, 2)  (R/predCrossVar.R:128-138)."""
    n = T * (T + 1) // 2
    a = np.empty(n, np.int32)
    b = np.empty(n, np.int32)
    _capi.lib().gms_trait_pairs(T, _ptr(a), _ptr(b))
    return a, b


class CrossVarEngine:
    def __init__(self, device=0):
        self._lib = _capi.lib()
        h = C.c_void_p()
        rc = self._lib.gms_create(C.byref(h), int(device))
        if rc != 0:
            raise GmsError(rc, self._lib.gms_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self.m = None
        self.P = None
        self._keep = []

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise GmsError(rc, self._lib.gms_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gms_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream):
        self._ck(self._lib.gms_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else None)))

    # ------------------------------------------------------------------ stage (1)
    def set_genmap(self, cM, m_c):
        """Genetic map (cM, chromosome-major) + SNPs per chromosome; the device builds R and RoR."""
        cM = np.ascontiguousarray(cM, dtype=np.float64)
        m_c = np.ascontiguousarray(m_c, dtype=np.int64)
        if int(m_c.sum()) != cM.size:
            raise ValueError("sum(m_c) != len(cM)")
        self._ck(self._lib.gms_set_genmap(self._h, int(m_c.size), _ptr(m_c), _ptr(cM)))
        self.m = int(cM.size)
        self.m_c = m_c.copy()
        self.P = None

    def set_recomb_blocks(self, blocks):
        """Diagonal blocks of recombFreqMat (= 1-2c), one square array per chromosome."""
        blocks = [np.asfortranarray(b, dtype=np.float64) for b in blocks]
        for b in blocks:
            if b.ndim != 2 or b.shape[0] != b.shape[1]:
                raise ValueError("blocks must be square")
        m_c = np.array([b.shape[0] for b in blocks], dtype=np.int64)
        ptrs = (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])
        self._ck(self._lib.gms_set_recomb_blocks(self._h, len(blocks), _ptr(m_c), C.cast(ptrs, C.c_void_p)))
        self.m = int(m_c.sum())
        self.m_c = m_c
        self.P = None

    def set_recomb_matrix(self, R, chrom):
        """Dense m x m recombFreqMat with SNPs already grouped by chromosome (`chrom[j]` labels).
        Uses per-chromosome blocks if every off-chromosome entry is exactly 0 (always true for
        1-2*genmap2recombfreq()), else treats the genome as one block."""
        R = np.asarray(R, dtype=np.float64)
        chrom = np.asarray(chrom)
        m = R.shape[0]
        if R.shape != (m, m) or chrom.shape != (m,):
            raise ValueError("recombFreqMat must be m x m and chrom length m")
        bounds = [0] + [i for i in range(1, m) if chrom[i] != chrom[i - 1]] + [m]
        if len(set(chrom[b] for b in bounds[:-1])) != len(bounds) - 1:
            raise ValueError("SNPs must be grouped by chromosome")
        mask_ok = True
        for a, b in zip(bounds[:-1], bounds[1:]):
            if np.any(R[a:b, :a] != 0.0) or np.any(R[a:b, b:] != 0.0):
                mask_ok = False
                break
        if mask_ok:
            self.set_recomb_blocks([R[a:b, a:b] for a, b in zip(bounds[:-1], bounds[1:])])
        else:
            self.set_recomb_blocks([R])
        return mask_ok

    def set_mode(self, mode):
        """"auto" (default: "int8_tensor" whenever its operand images fit in HBM, else "fp64"), "fp64" / "dense" (FP64
        DMMA contraction), "int8_tensor" (int8 digit planes on tcgen05 with an a-posteriori error bound and FP64
        re-evaluation of the units that fail it), "map_scan" (map-structured prefix-sum fast path; needs set_genmap()
        with sorted positions)."""
        code = {"dense": 0, "fp64": 0, "map_scan": 1, "int8_tensor": 2, "auto": 3}[mode] if isinstance(mode, str) else int(mode)
        self._ck(self._lib.gms_set_mode(self._h, code))

    def set_int8_tolerance(self, tau):
        """tau of the int8 error bound (default 2**-33); <= 0 switches the verification off."""
        self._ck(self._lib.gms_set_int8_tolerance(self._h, float(tau)))

    def set_cache(self, enabled):
        """Per-handle cache of operand images and per-parent table rows, keyed on a hash of the effects (default on)."""
        self._ck(self._lib.gms_set_cache(self._h, 1 if enabled else 0))

    def set_parent_shard(self, index, parts):
        """This rank evaluates slice `index` of `parts` of the per-parent stage (see stage_range / exchange_parents)."""
        self._ck(self._lib.gms_set_parent_shard(self._h, int(index), int(parts)))

    def set_chromosome_shard(self, c_begin, c_end):
        self._ck(self._lib.gms_set_chromosome_shard(self._h, int(c_begin), int(c_end)))

    # ------------------------------------------------------------------ haplotypes
    def set_haplotypes(self, hap):
        """(2P) x m matrix in {0,1}; parent p = rows 2p (HapA), 2p+1 (HapB).
        int32 Fortran-ordered input goes through the R-layout entry point, anything else as uint8."""
        hap = np.asarray(hap)
        if hap.ndim != 2 or hap.shape[0] % 2:
            raise ValueError("haplotype matrix must be (2P) x m")
        P, m = hap.shape[0] // 2, hap.shape[1]
        if hap.dtype == np.int32 and hap.flags.f_contiguous and not hap.flags.c_contiguous:
            self._ck(self._lib.gms_set_haplotypes_i32cm(self._h, P, m, _ptr(hap)))
        else:
            if hap.dtype != np.uint8:
                if np.any((hap != 0) & (hap != 1)):
                    raise GmsError(_capi.GMS_EINVAL, "haploMat has values outside {0,1}")
                hap = hap.astype(np.uint8)
            hap = np.ascontiguousarray(hap)
            self._ck(self._lib.gms_set_haplotypes_u8rm(self._h, P, m, _ptr(hap)))
        self.P = P

    # ------------------------------------------------------------------ the hot path
    @staticmethod
    def _eff(a, m):
        if a is None:
            return None, None
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        if a.shape[0] != m:
            raise ValueError(f"effects must be m x T with m = {m}, got {a.shape}")
        return np.asfortranarray(a), a.shape[1]

    def stage(self, model, add, dom=None, asub=None, sire=None, dam=None):
        model = MODEL[model] if isinstance(model, str) else int(model)
        add, T = self._eff(add, self.m)
        dom, Td = self._eff(dom, self.m)
        asub, Ta = self._eff(asub, self.m)
        for t in (Td, Ta):
            if t is not None and t != T:
                raise ValueError("effect matrices must have the same number of traits")
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        if sire.shape != dam.shape or sire.ndim != 1:
            raise ValueError("sire and dam must be 1-D and of equal length")
        self._ck(self._lib.gms_stage_inputs(self._h, model, T, _ptr(add), _ptr(dom), _ptr(asub),
                                            sire.size, _ptr(sire), _ptr(dam)))
        self._shape = (sire.size, T * (T + 1) // 2, model + 1)
        return self._shape

    def stage_range(self, model, add, dom=None, asub=None, sire=None, dam=None, lo=0, hi=None):
        """gms_stage_inputs_range: the WHOLE cross list plus this rank's range [lo, hi) of it."""
        model = MODEL[model] if isinstance(model, str) else int(model)
        add, T = self._eff(add, self.m)
        dom, _ = self._eff(dom, self.m)
        asub, _ = self._eff(asub, self.m)
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        hi = sire.size if hi is None else int(hi)
        self._ck(self._lib.gms_stage_inputs_range(self._h, model, T, _ptr(add), _ptr(dom), _ptr(asub), sire.size,
                                                  _ptr(sire), _ptr(dam), int(lo), hi))
        self._shape = (hi - int(lo), T * (T + 1) // 2, model + 1)
        return self._shape

    def compute(self, sync=True):
        self._ck(self._lib.gms_compute(self._h, 1 if sync else 0))

    def compute_parents(self):
        self._ck(self._lib.gms_compute_parents(self._h))

    def parent_exchange(self):
        """(device pointer, doubles per part, parts) of the exchange buffer of a parent-sharded handle; pointer 0 when
        no table row had to be computed (all cached)."""
        ptr, n, parts = C.c_void_p(), C.c_int64(), C.c_int32()
        self._ck(self._lib.gms_parent_exchange(self._h, C.byref(ptr), C.byref(n), C.byref(parts)))
        return (ptr.value or 0), int(n.value), int(parts.value)

    def exchange_parents(self, dist, group=None):
        """One all-gather (NCCL) of the per-parent table slices over `group`, in place in the engine's buffer."""
        import torch
        ptr, n, parts = self.parent_exchange()
        if not ptr or parts <= 1:
            return

        class _Buf:        # a zero-copy torch view of the library's device buffer
            __cuda_array_interface__ = {"shape": (parts * n,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                        "strides": None}
        full = torch.as_tensor(_Buf(), device=torch.device("cuda", self.device))
        rank = dist.get_rank(group)
        dist.all_gather_into_tensor(full, full[rank * n:(rank + 1) * n], group=group)

    def device_outputs(self):
        """Zero-copy torch views (float64 [N, npairs, ncomp], int32 [N]) of the result slab and Nsegsnps in HBM."""
        import torch
        o, s, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self._lib.gms_device_outputs(self._h, C.byref(o), C.byref(s), C.byref(n)))
        N, npairs, ncomp = self._shape
        dev = torch.device("cuda", self.device)
        if not n.value:
            return torch.empty((0, npairs, ncomp), dtype=torch.float64, device=dev), torch.empty(0, dtype=torch.int32, device=dev)

        def view(ptr, shape, typestr):
            class _Buf:
                __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}
            return torch.as_tensor(_Buf(), device=dev)
        return view(o.value, (N, npairs, ncomp), "<f8"), view(s.value, (N,), "<i4")

    def compute_crosses(self, sync=True):
        self._ck(self._lib.gms_compute_crosses(self._h, 1 if sync else 0))

    def fetch(self, out=None, nseg=None):
        N, npairs, ncomp = self._shape
        if out is None:
            out = np.empty((N, npairs, ncomp), np.float64)
        if nseg is None:
            nseg = np.empty(N, np.int32)
        self._ck(self._lib.gms_fetch_outputs(self._h, _ptr(out), _ptr(nseg)))
        return out, nseg

    def predict(self, model, add, dom=None, asub=None, sire=None, dam=None, out=None, nseg=None):
        """One gms_predict call: host buffers in, host buffers out.
        Returns (out[N, npairs, ncomp], Nsegsnps[N])."""
        model = MODEL[model] if isinstance(model, str) else int(model)
        add, T = self._eff(add, self.m)
        dom, Td = self._eff(dom, self.m)
        asub, Ta = self._eff(asub, self.m)
        for t in (Td, Ta):
            if t is not None and t != T:
                raise ValueError("effect matrices must have the same number of traits")
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        if sire.shape != dam.shape or sire.ndim != 1:
            raise ValueError("sire and dam must be 1-D and of equal length")
        N, npairs, ncomp = sire.size, T * (T + 1) // 2, model + 1
        if out is None:
            out = np.empty((N, npairs, ncomp), np.float64)
        if nseg is None:
            nseg = np.empty(N, np.int32)
        self._shape = (N, npairs, ncomp)
        self._ck(self._lib.gms_predict(self._h, model, T, _ptr(add), _ptr(dom), _ptr(asub), N,
                                       _ptr(sire), _ptr(dam), _ptr(out), _ptr(nseg)))
        return out, nseg

    def predict_batched(self, model, add_sets, dom_sets=None, asub_sets=None, sire=None, dam=None):
        """gms_predict_batched: several effect sets (cross-validation folds) against one cross list.
        add_sets: [nsets, m, T]. Returns (out[nsets, N, npairs, ncomp], Nsegsnps[N])."""
        model = MODEL[model] if isinstance(model, str) else int(model)

        def pack(sets):
            if sets is None:
                return None
            a = np.asarray(sets, dtype=np.float64)
            if a.ndim != 3 or a.shape[1] != self.m:
                raise ValueError("effect sets must be [nsets, m, T]")
            return np.ascontiguousarray(a.transpose(0, 2, 1))        # per set: m x T column-major
        add, dom, asub = pack(add_sets), pack(dom_sets), pack(asub_sets)
        nsets, T = add.shape[0], add.shape[1]
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        N, npairs, ncomp = sire.size, T * (T + 1) // 2, model + 1
        out = np.empty((nsets, N, npairs, ncomp), np.float64)
        nseg = np.empty(N, np.int32)
        self._ck(self._lib.gms_predict_batched(self._h, model, nsets, T, _ptr(add), _ptr(dom), _ptr(asub), N,
                                               _ptr(sire), _ptr(dam), _ptr(out), _ptr(nseg)))
        return out, nseg

    @staticmethod
    def predict_sharded(engines, model, add, dom=None, asub=None, sire=None, dam=None):
        """gms_predict_sharded: several engines (one per GPU, same map + haplotypes) in ONE process; the
        cross list is cut into contiguous shards, results land in one host slab."""
        e0 = engines[0]
        model = MODEL[model] if isinstance(model, str) else int(model)
        add, T = e0._eff(add, e0.m)
        dom, _ = e0._eff(dom, e0.m)
        asub, _ = e0._eff(asub, e0.m)
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        N, npairs, ncomp = sire.size, T * (T + 1) // 2, model + 1
        out = np.empty((N, npairs, ncomp), np.float64)
        nseg = np.empty(N, np.int32)
        hs = (C.c_void_p * len(engines))(*[e._h for e in engines])
        rc = e0._lib.gms_predict_sharded(C.cast(hs, C.c_void_p), len(engines), model, T, _ptr(add), _ptr(dom), _ptr(asub),
                                         N, _ptr(sire), _ptr(dam), _ptr(out), _ptr(nseg))
        e0._ck(rc)
        return out, nseg

    # ------------------------------------------------------------------ "next" rows
    def cross_means(self, add, dom=None, sire=None, dam=None):
        """predCrossMeans: returns (gebv[P,T], mean_bv[N,T], mean_tgv[N,T] or None)."""
        add, T = self._eff(add, self.m)
        dom, _ = self._eff(dom, self.m)
        sire = np.ascontiguousarray(sire, dtype=np.int32)
        dam = np.ascontiguousarray(dam, dtype=np.int32)
        gebv = np.empty((self.P, T), np.float64)
        mbv = np.empty((sire.size, T), np.float64)
        mtgv = np.empty((sire.size, T), np.float64) if dom is not None else None
        self._ck(self._lib.gms_cross_means(self._h, T, _ptr(add), _ptr(dom), sire.size, _ptr(sire), _ptr(dam),
                                           _ptr(gebv), _ptr(mbv), _ptr(mtgv)))
        return gebv, mbv, mtgv

    def tidy(self, si_weights=None, std_sel_int=2.67):
        """gms_tidy: predictCrosses()' tidy table for the last predict / compute, computed on the device.
        Returns tidy[N, nof, nslots, 4] = (predMean, predVar, predSD, predUsefulness); predOf axis: BV [, TGV];
        slot axis: [SELIND,] traits."""
        N, npairs, ncomp = self._shape
        T = int((np.sqrt(8 * npairs + 1) - 1) / 2)
        w = None if si_weights is None else np.ascontiguousarray(si_weights, dtype=np.float64)
        if w is not None and w.size != T:
            raise ValueError("si_weights must have one weight per trait")
        out = np.empty((N, 1 if ncomp == 1 else 2, T + (w is not None), 4), np.float64)
        self._ck(self._lib.gms_tidy(self._h, _ptr(w), float(std_sel_int), _ptr(out)))
        return out

    @staticmethod
    def selind(model, out, w):
        """Selection-index variance per cross and component (R/gmsFunctions.R:825-853)."""
        model = MODEL[model] if isinstance(model, str) else int(model)
        out = np.ascontiguousarray(out, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        N, npairs, ncomp = out.shape
        T = w.size
        assert npairs == T * (T + 1) // 2 and ncomp == model + 1
        si = np.empty((N, ncomp), np.float64)
        rc = _capi.lib().gms_selind(model, T, N, _ptr(out), _ptr(w), _ptr(si))
        if rc != 0:
            raise GmsError(rc, "gms_selind: bad arguments")
        return si

    def stats(self):
        s = _capi.Stats()
        self._ck(self._lib.gms_get_stats(self._h, C.byref(s)))
        return s.asdict()
