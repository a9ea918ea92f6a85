#!/bin/bash
# Dev tool (GPU box or any host): bash tools/lis_bench/run.sh -- builds and times the variants of the host LIS.
set -e
cd "$(dirname "$0")"
ROOT=../..
python - <<'PY'
import sys
sys.path.insert(0, "../..")
from miniwfa_b200 import synth
t, q = synth.make_pair(5000000, 0.03, 424242)
open("t.bin", "wb").write(t); open("q.bin", "wb").write(q)
PY
rm -f hits.bin
for A in 16 32 64 128; do
  sed "s/#define LIS_AHEAD 16/#define LIS_AHEAD $A/" $ROOT/miniwfa_b200/csrc/mwf_chain.c > chain_$A.c
  gcc -O2 -DLIS2 -DCHAIN_C="\"chain_$A.c\"" -I $ROOT/include -I $ROOT/miniwfa_b200/csrc -o bench_$A bench.c $ROOT/miniwfa_b200/csrc/kalloc.c 2>/dev/null
  ./bench_$A "prefetch distance $A"
done
rm -f chain_*.c bench_16 bench_32 bench_64 bench_128 t.bin q.bin hits.bin
