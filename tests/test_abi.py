"""CPU: the C-ABI library builds, loads, exports every symbol the headers declare, keeps the reference's
struct layout, and its host-only parts (kalloc, CIGAR helpers) behave.  No GPU compute is called."""
import ctypes
import os
import re

from conftest import ROOT
from miniwfa_b200 import api


def declared_functions():
    names = set()
    for h in ("miniwfa.h", "kalloc.h", "mwf_b200.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"#define.*?(?<!\\)\n", "\n", src, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src):
            names.add(m.group(1))
    return names


def test_exports_every_declared_symbol(product_lib):
    names = declared_functions()
    assert {"mwf_opt_init", "mwf_wfa_exact", "mwf_wfa_auto", "mwf_wfa_chain", "mwf_cigar2score", "mwf_assert_cigar",
            "kmalloc", "kcalloc", "krealloc", "krelocate", "kfree", "km_init", "km_init2", "km_destroy", "km_stat",
            "mwf_wfa_exact_batch", "mwf_b200_batch_create", "mwf_b200_batch_run"} <= names
    for n in sorted(names):
        assert hasattr(product_lib, n), "missing export: " + n


def test_struct_layout_matches_reference():
    assert ctypes.sizeof(api.MwfOpt) == 56 and ctypes.sizeof(api.MwfRst) == 24
    assert api.MwfOpt.max_iter.offset == 32 and api.MwfOpt.max_occ.offset == 40
    assert api.MwfRst.n_iter.offset == 8 and api.MwfRst.cigar.offset == 16


def test_opt_init_defaults(product_lib):
    o = api.opt_init()
    assert (o.flag, o.x, o.o1, o.e1, o.o2, o.e2, o.step, o.max_s, o.max_iter) == (0, 4, 4, 2, 15, 1, 0, 0, 0)
    assert (o.kmer, o.max_occ, o.min_len) == (13, 2, 30)


def test_cigar2score(product_lib):
    o = api.opt_init()
    cig = [1 << 4 | 8, 16 << 4 | 7, 1 << 4 | 8, 14 << 4 | 7, 128 << 4 | 1, 4 << 4 | 7, 1 << 4 | 8, 24 << 4 | 7]
    assert api.cigar_string(cig) == "1X16=1X14=128I4=1X24="
    assert api.cigar2score(o, cig) == (155, 61, 189)


def test_kalloc_arena(product_lib):
    L = product_lib
    L.km_init2.restype = ctypes.c_void_p
    L.km_init2.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.kcalloc.restype = ctypes.c_void_p
    L.kcalloc.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
    L.krealloc.restype = ctypes.c_void_p
    L.krealloc.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.krelocate.restype = ctypes.c_void_p
    L.krelocate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]

    class Stat(ctypes.Structure):
        _fields_ = [(n, ctypes.c_size_t) for n in ("capacity", "available", "n_blocks", "n_cores", "largest")]
    L.km_stat.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stat)]

    km = L.km_init()
    assert km
    assert L.kmalloc(km, 0) is None
    import random
    rng = random.Random(1)
    live = {}
    for it in range(3000):
        if live and rng.random() < 0.45:
            p = rng.choice(list(live))
            n, tag = live.pop(p)
            buf = (ctypes.c_ubyte * n).from_address(p)
            assert all(b == tag for b in buf[:min(n, 64)]) and buf[n - 1] == tag
            L.kfree(km, p)
        else:
            n = rng.choice([1, 7, 16, 100, 1000, 70000, 9000000 if it % 500 == 0 else 33])
            p = L.kcalloc(km, n, 1)
            assert p and p % 16 == 0 and p not in live
            buf = (ctypes.c_ubyte * n).from_address(p)
            assert buf[0] == 0 and buf[n - 1] == 0
            tag = it % 251 + 1
            ctypes.memset(p, tag, n)
            live[p] = (n, tag)
    # grow / relocate keep contents
    p = L.kmalloc(km, 10)
    ctypes.memmove(p, b"0123456789", 10)
    p2 = L.krealloc(km, p, 100000)
    assert ctypes.string_at(p2, 10) == b"0123456789"
    assert L.krealloc(km, p2, 5) == p2  # never shrinks
    p3 = L.krelocate(km, p2, 10)
    assert ctypes.string_at(p3, 10) == b"0123456789"
    # child arena: allocations come out of the parent and go back on destroy
    st0, st1 = Stat(), Stat()
    L.km_stat(km, ctypes.byref(st0))
    child = L.km_init2(km, 0)
    q = L.kmalloc(child, 5000)
    ctypes.memset(q, 7, 5000)
    L.km_destroy(child)
    for p_ in list(live):
        L.kfree(km, p_)
    L.kfree(km, p3)
    L.km_stat(km, ctypes.byref(st1))
    assert st1.capacity >= st0.capacity > 0 and st1.n_cores >= 1
    assert st1.available >= st1.capacity - 64 * st1.n_cores  # everything returned, free list coalesced
    assert st1.n_blocks <= st1.n_cores
    L.km_destroy(km)
    # NULL arena forwards to libc
    p = L.kmalloc(None, 32)
    assert p
    L.kfree(None, p)
    assert L.krelocate(None, 1234, 8) == 1234


def test_no_cpu_fallback_symbols(product_lib):
    """The product library must not embed the oracle."""
    out = os.popen("nm -D %s" % api.LIB_PATH).read()
    assert "orc_" not in out


def test_reference_cli_links_against_the_library(product_lib):
    """Where /root/reference exists: its unmodified main.c compiles against include/miniwfa.h and links with the product library
    in place of miniwfa.o kalloc.o mwf-dbg.o (the drop-in of INTEGRATION.md 1); every API symbol it needs resolves to us."""
    import shutil
    import subprocess
    import pytest
    if not os.path.exists("/root/reference/main.c"):
        pytest.skip("/root/reference is not present on this machine")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    exe = os.path.join(ROOT, "oracle", "_ref", "test-mwf-dropin")
    assert os.path.exists(exe)
    und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True, check=True).stdout
    assert {"mwf_opt_init", "mwf_wfa_exact", "mwf_wfa_chain", "mwf_wfa_auto", "mwf_assert_cigar"} <= {ln.split()[-1] for ln in und.splitlines() if ln.split()}
    if shutil.which("ldd"):
        out = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
        assert "libminiwfa_b200.so" in out and "not found" not in out.split("libminiwfa_b200.so")[1].split("\n")[0]


def test_kalloc_pool_macros(product_lib, tmp_path):
    """KALLOC_POOL_INIT (reference kalloc.h:43-80): a C program using the typed pool compiles against include/kalloc.h, links
    with the product library, and recycles freed objects before asking the arena."""
    import subprocess
    src = tmp_path / "pool.c"
    src.write_text(r'''
#include <stdio.h>
#include "kalloc.h"
typedef struct { int a; double b; } node_t;
KALLOC_POOL_INIT(node, node_t)
int main(void)
{
	void *km = km_init();
	kmp_node_t *mp = kmp_init_node(km);
	node_t *x[40], *y;
	int i, fresh_zero = 1, reused = 0;
	for (i = 0; i < 40; ++i) { x[i] = kmp_alloc_node(mp); fresh_zero &= x[i]->a == 0 && x[i]->b == 0.0; x[i]->a = i + 1; }
	for (i = 0; i < 40; ++i) kmp_free_node(mp, x[i]);
	y = kmp_alloc_node(mp);
	for (i = 0; i < 40; ++i) reused |= y == x[i];
	printf("%d %d %d %d %d\n", fresh_zero, reused, (int)mp->cnt, (int)mp->n, y->a);
	kmp_free_node(mp, y);
	kmp_destroy_node(mp);
	km_destroy(km);
	return 0;
}
''')
    exe = tmp_path / "pool"
    subprocess.run(["gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", os.path.dirname(api.LIB_PATH), "-lminiwfa_b200", "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["1", "1", "1", "39", "40"], out
