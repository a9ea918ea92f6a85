"""TEST INFRASTRUCTURE: ctypes loaders for the CPU checkers (never imported by the product package).

  oracle()    -> liboracle.so, the C restatement in wfa_oracle.c (built on demand with gcc)
  reference() -> _ref/libminiwfa_ref.so, the unmodified reference, or None when it was never built
Both take / fill the same struct layouts as the product (mwf_opt_t / mwf_rst_t).
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


class Opt(ctypes.Structure):
    _fields_ = [("flag", ctypes.c_int32), ("x", ctypes.c_int32), ("o1", ctypes.c_int32), ("e1", ctypes.c_int32),
                ("o2", ctypes.c_int32), ("e2", ctypes.c_int32), ("step", ctypes.c_int32), ("max_s", ctypes.c_int32),
                ("max_iter", ctypes.c_int64), ("max_occ", ctypes.c_int32), ("kmer", ctypes.c_int32),
                ("min_len", ctypes.c_int32)]


class Rst(ctypes.Structure):
    _fields_ = [("s", ctypes.c_int32), ("n_cigar", ctypes.c_int32), ("n_iter", ctypes.c_int64),
                ("cigar", ctypes.POINTER(ctypes.c_uint32))]


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference exists, _ref/."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)


_orc = None
_ref = None


def oracle():
    global _orc
    if _orc is None:
        path = os.path.join(HERE, "liboracle.so")
        src = os.path.join(HERE, "wfa_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        L.orc_wfa_exact.argtypes = [ctypes.POINTER(Opt), ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                    ctypes.c_char_p, ctypes.POINTER(Rst)]
        L.orc_wfa_auto_exact_leg.argtypes = L.orc_wfa_exact.argtypes
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_wfa_checkpoints.restype = ctypes.POINTER(ctypes.c_int32)
        L.orc_wfa_checkpoints.argtypes = [ctypes.POINTER(Opt), ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                                          ctypes.c_char_p, ctypes.POINTER(ctypes.c_int32)]
        _orc = L
    return _orc


def reference():
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libminiwfa_ref.so")
        if not os.path.exists(path):
            return None
        L = ctypes.CDLL(path)
        for f in (L.mwf_wfa_exact, L.mwf_wfa_auto, L.mwf_wfa_chain):
            f.argtypes = [ctypes.c_void_p, ctypes.POINTER(Opt), ctypes.c_int32, ctypes.c_char_p, ctypes.c_int32,
                          ctypes.c_char_p, ctypes.POINTER(Rst)]
            f.restype = None
        _ref = L
    return _ref


_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]


def make_opt(**kw):
    o = Opt()
    oracle().orc_opt_init(ctypes.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def copy_opt(src):
    """Opt from any struct with the same fields (e.g. the product's MwfOpt)."""
    o = Opt()
    for name, _ in Opt._fields_:
        setattr(o, name, getattr(src, name))
    return o


def _take(r, free):
    cig = [r.cigar[i] for i in range(r.n_cigar)] if r.n_cigar > 0 else []
    if r.cigar:
        free(r.cigar)
    return (r.s, r.n_cigar, r.n_iter, cig)


def oracle_exact(opt, ts, qs):
    r = Rst()
    oracle().orc_wfa_exact(ctypes.byref(opt), len(ts), ts, len(qs), qs, ctypes.byref(r))
    return _take(r, oracle().orc_free)


def reference_exact(opt, ts, qs, fn="mwf_wfa_exact"):
    L = reference()
    if L is None:
        raise RuntimeError("oracle/_ref has not been built")
    r = Rst()
    getattr(L, fn)(None, ctypes.byref(opt), len(ts), ts, len(qs), qs, ctypes.byref(r))
    return _take(r, _libc.free)


def checker_exact(opt, ts, qs):
    """The strongest checker available: the real reference when built, else the restatement."""
    if reference() is not None and not (len(ts) == 0 and len(qs) == 0 and (opt.flag & 1)):
        return reference_exact(opt, ts, qs)
    return oracle_exact(opt, ts, qs)
