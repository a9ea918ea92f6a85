"""ctypes mirror of include/miniwfa.h and include/mwf_b200.h (same names, same argument meaning).

Nothing here computes alignments: every call goes through the C-ABI of libminiwfa_b200.so,
which runs the CUDA kernels.  If the library is absent this module raises at first use.
"""
import ctypes
import os

import numpy as _np

F_CIGAR = 0x1
F_NO_KALLOC = 0x2
KERNEL_AUTO, KERNEL_CTA, KERNEL_GRID, KERNEL_TILE = 0, 1, 2, 3

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libminiwfa_b200.so")


class MwfOpt(ctypes.Structure):  # mwf_opt_t, 56 bytes
    _fields_ = [("flag", ctypes.c_int32), ("x", ctypes.c_int32), ("o1", ctypes.c_int32), ("e1", ctypes.c_int32),
                ("o2", ctypes.c_int32), ("e2", ctypes.c_int32), ("step", ctypes.c_int32), ("max_s", ctypes.c_int32),
                ("max_iter", ctypes.c_int64), ("max_occ", ctypes.c_int32), ("kmer", ctypes.c_int32),
                ("min_len", ctypes.c_int32)]


class MwfRst(ctypes.Structure):  # mwf_rst_t, 24 bytes
    _fields_ = [("s", ctypes.c_int32), ("n_cigar", ctypes.c_int32), ("n_iter", ctypes.c_int64),
                ("cigar", ctypes.POINTER(ctypes.c_uint32))]


_lib = None


def lib():
    """The loaded product library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libminiwfa_b200.so is missing: run `python -m miniwfa_b200.build` "
                               "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        pO, pR = ctypes.POINTER(MwfOpt), ctypes.POINTER(MwfRst)
        L.mwf_opt_init.argtypes = [pO]
        for f in (L.mwf_wfa_exact, L.mwf_wfa_auto, L.mwf_wfa_chain):
            f.argtypes = [vp, pO, i32, ctypes.c_char_p, i32, ctypes.c_char_p, pR]
            f.restype = None
        L.mwf_cigar2score.argtypes = [pO, i32, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.mwf_cigar2score.restype = i32
        L.mwf_wfa_exact_batch.argtypes = [vp, pO, i32, ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_char_p), pR]
        L.mwf_b200_batch_create.argtypes = [pO, i32, ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.mwf_b200_batch_create.restype = vp
        L.mwf_b200_batch_set_stream.argtypes = [vp, vp]
        L.mwf_b200_batch_upload.argtypes = [vp, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p)]
        for name in ("run", "wait", "destroy"):
            getattr(L, "mwf_b200_batch_" + name).argtypes = [vp]
            getattr(L, "mwf_b200_batch_" + name).restype = None
        L.mwf_b200_batch_fetch.argtypes = [vp, vp, pR]
        L.mwf_b200_batch_kernel_ms.argtypes = [vp]
        L.mwf_b200_batch_kernel_ms.restype = ctypes.c_double
        for name in ("launches", "h2d_bytes", "d2h_bytes"):
            getattr(L, "mwf_b200_batch_" + name).argtypes = [vp]
            getattr(L, "mwf_b200_batch_" + name).restype = i64
        L.mwf_b200_batch_kernel_used.argtypes = [vp]
        L.mwf_b200_batch_kernel_used.restype = ctypes.c_int
        L.mwf_b200_release_cache.argtypes = []
        L.mwf_b200_release_cache.restype = None
        L.mwf_b200_kmer_hits.argtypes = [i32, ctypes.c_char_p, i32, ctypes.c_char_p, i32, i32,
                                         ctypes.POINTER(ctypes.POINTER(ctypes.c_uint64))]
        L.mwf_b200_kmer_hits.restype = i64
        L.mwf_b200_kmer_free.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
        L.mwf_b200_kmer_free.restype = None
        L.mwf_b200_kmer_shared.argtypes = [i32, ctypes.c_char_p, i32, ctypes.c_char_p, i32,
                                           ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]
        L.mwf_b200_kmer_shared.restype = None
        L.mwf_b200_kmer_launches.argtypes = []
        L.mwf_b200_kmer_launches.restype = i64
        L.kfree.argtypes = [vp, vp]
        L.km_init.restype = vp
        L.km_destroy.argtypes = [vp]
        L.kmalloc.argtypes = [vp, ctypes.c_size_t]
        L.kmalloc.restype = vp
        _lib = L
    return _lib


def opt_init(**kw):
    """mwf_opt_init() plus keyword overrides (flag=, x=, o1=, ..., step=, max_s=, max_iter=)."""
    o = MwfOpt()
    lib().mwf_opt_init(ctypes.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def device_count():
    return lib().mwf_b200_device_count()


def set_device(dev):
    lib().mwf_b200_set_device(int(dev))


def set_devices(n):
    """mwf_b200_set_devices(): devices one mwf_wfa_exact_batch() call is spread over (0 = automatic, 1 = the current one)."""
    lib().mwf_b200_set_devices(int(n))


def set_kernel(kernel):
    lib().mwf_b200_set_kernel(int(kernel))


def release_cache():
    """mwf_b200_release_cache(): free every cached device / pinned-host workspace."""
    lib().mwf_b200_release_cache()


def _take(r, km=None):
    """Copy a result out of an mwf_rst_t and release its CIGAR (allocated from km / malloc)."""
    cig = _np.ctypeslib.as_array(r.cigar, shape=(r.n_cigar,)).tolist() if r.n_cigar > 0 else []  # one C loop, not one ctypes call per word
    if r.cigar:
        lib().kfree(km, r.cigar)
    return (r.s, r.n_cigar, r.n_iter, cig)


def wfa_exact(opt, ts, qs, km=None):
    """mwf_wfa_exact(km, opt, tl, ts, ql, qs, &r) -> (s, n_cigar, n_iter, [cigar words])."""
    r = MwfRst()
    lib().mwf_wfa_exact(km, ctypes.byref(opt), len(ts), ts, len(qs), qs, ctypes.byref(r))
    return _take(r, km)


def wfa_auto(opt, ts, qs, km=None):
    r = MwfRst()
    lib().mwf_wfa_auto(km, ctypes.byref(opt), len(ts), ts, len(qs), qs, ctypes.byref(r))
    return _take(r, km)


def wfa_chain(opt, ts, qs, km=None):
    r = MwfRst()
    lib().mwf_wfa_chain(km, ctypes.byref(opt), len(ts), ts, len(qs), qs, ctypes.byref(r))
    return _take(r, km)


def kmer_hits(ts, qs, k, max_occ):
    """mwf_b200_kmer_hits -> list of query << 32 | target words in ascending (target, query) order."""
    p = ctypes.POINTER(ctypes.c_uint64)()
    n = lib().mwf_b200_kmer_hits(len(ts), ts, len(qs), qs, k, max_occ, ctypes.byref(p))
    out = [p[i] for i in range(n)]
    lib().mwf_b200_kmer_free(p)
    return out


def kmer_shared(s1, s2, k):
    """mwf_b200_kmer_shared -> (k-mers in s1, k-mers in s2, sum of min(copies) over distinct k-mers)."""
    a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    lib().mwf_b200_kmer_shared(len(s1), s1, len(s2), s2, k, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def _arrays(pairs):
    n = len(pairs)
    tl = (ctypes.c_int32 * n)(*[len(p[0]) for p in pairs])
    ql = (ctypes.c_int32 * n)(*[len(p[1]) for p in pairs])
    ts = (ctypes.c_char_p * n)(*[bytes(p[0]) for p in pairs])
    qs = (ctypes.c_char_p * n)(*[bytes(p[1]) for p in pairs])
    return n, tl, ts, ql, qs


def host_arrays(pairs):
    """The C arrays (n, tl[], ts[], ql[], qs[]) mwf_wfa_exact_batch / mwf_b200_batch_* take, built once."""
    return _arrays(pairs)


def wfa_exact_batch(opt, pairs, km=None):
    """mwf_wfa_exact_batch over [(ts, qs), ...] -> list of (s, n_cigar, n_iter, [cigar words])."""
    n, tl, ts, ql, qs = _arrays(pairs)
    r = (MwfRst * n)()
    lib().mwf_wfa_exact_batch(km, ctypes.byref(opt), n, tl, ts, ql, qs, r)
    return [_take(r[i], km) for i in range(n)]


class Batch:
    """mwf_b200_batch_*: create -> upload -> run -> wait -> fetch, with the engine's own timers."""

    def __init__(self, opt, pairs, arrays=None):
        """arrays: the host buffers of `pairs` as returned by host_arrays(pairs), to build them only once."""
        self.n, self._tl, self._ts, self._ql, self._qs = arrays if arrays is not None else _arrays(pairs)
        self.opt = opt
        self.h = lib().mwf_b200_batch_create(ctypes.byref(opt), self.n, self._tl, self._ql)

    def set_stream(self, cuda_stream):
        lib().mwf_b200_batch_set_stream(self.h, cuda_stream)

    def upload(self):
        lib().mwf_b200_batch_upload(self.h, self._ts, self._qs)

    def run(self):
        lib().mwf_b200_batch_run(self.h)

    def wait(self):
        lib().mwf_b200_batch_wait(self.h)

    def fetch(self, km=None):
        r = (MwfRst * self.n)()
        lib().mwf_b200_batch_fetch(self.h, km, r)
        return [_take(r[i], km) for i in range(self.n)]

    @property
    def kernel_ms(self):
        return lib().mwf_b200_batch_kernel_ms(self.h)

    @property
    def launches(self):
        return lib().mwf_b200_batch_launches(self.h)

    @property
    def kernel_used(self):
        return lib().mwf_b200_batch_kernel_used(self.h)

    @property
    def h2d_bytes(self):
        return lib().mwf_b200_batch_h2d_bytes(self.h)

    @property
    def d2h_bytes(self):
        return lib().mwf_b200_batch_d2h_bytes(self.h)

    def close(self):
        if self.h:
            lib().mwf_b200_batch_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def cigar_string(cig):
    return "".join("%d%s" % (c >> 4, "MIDNSHP=XBid"[c & 0xf]) for c in cig)


def cigar2score(opt, cig):
    """mwf_cigar2score -> (score, target bases consumed, query bases consumed)."""
    n = len(cig)
    arr = (ctypes.c_uint32 * max(1, n))(*cig)
    tl, ql = ctypes.c_int32(), ctypes.c_int32()
    s = lib().mwf_cigar2score(ctypes.byref(opt), n, arr, ctypes.byref(tl), ctypes.byref(ql))
    return s, tl.value, ql.value
