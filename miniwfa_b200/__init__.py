"""miniwfa_b200 -- B200-native exact dual-affine wavefront aligner behind the miniwfa C API.

The product is the C-ABI library ``libminiwfa_b200.so`` (host C driver + sm_100a CUDA engine).
This package is thin plumbing over it: a ctypes mirror of ``miniwfa.h`` / ``mwf_b200.h`` for
tests and the benchmark, the synthetic pair generator, and the in-tree build script.
There is no CPU implementation here; loading fails loudly if the library is missing.
"""
from .api import (MwfOpt, MwfRst, Batch, opt_init, wfa_exact, wfa_auto, wfa_chain, wfa_exact_batch,
                  kmer_hits, kmer_shared,
                  cigar_string, cigar2score, device_count, set_device, set_devices, set_kernel, release_cache, lib,
                  F_CIGAR, F_NO_KALLOC, KERNEL_AUTO, KERNEL_CTA, KERNEL_GRID, KERNEL_TILE)
from . import synth

__all__ = ["MwfOpt", "MwfRst", "Batch", "opt_init", "wfa_exact", "wfa_auto", "wfa_exact_batch", "wfa_chain", "kmer_hits", "kmer_shared",
           "cigar_string", "cigar2score", "device_count", "set_device", "set_devices", "set_kernel", "release_cache", "lib", "synth",
           "F_CIGAR", "F_NO_KALLOC", "KERNEL_AUTO", "KERNEL_CTA", "KERNEL_GRID", "KERNEL_TILE"]
