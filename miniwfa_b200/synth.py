"""Deterministic synthetic sequence pairs: a counter-based splitmix64 generator, so any
language can reproduce pair i from its seed alone.

Target: n i.i.d. uniform bases over ACGT.  Query: every target position starts an event with probability p --
80 % substitution (one of the 3 other bases), 10 % insertion of L random bases before the position,
10 % deletion of L positions, L ~ Geometric(1/2) (mean 2).  Positions swallowed by a deletion carry no event.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_BASE = 0x5EED0000


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _stream(seed, stream, n):
    """n 64-bit values of the (seed, stream) sequence."""
    with np.errstate(over="ignore"):
        key = _splitmix64(np.array([(seed * 0x100 + stream) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
        return _splitmix64(key + np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95))


_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_MAX_INS = 32


def make_pair(n, p, index=0):
    """Return (target_bytes, query_bytes) for pair `index` (seed = SEED_BASE + index)."""
    seed = SEED_BASE + int(index)
    if n == 0:
        return b"", b""
    with np.errstate(over="ignore"):
        tcode = (_stream(seed, 0, n) >> np.uint64(62)).astype(np.int64)
        ev = (_stream(seed, 1, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) < p
        kind = (_stream(seed, 2, n) % np.uint64(10)).astype(np.int64)  # 0..7 sub, 8 ins, 9 del
        r3 = _stream(seed, 3, n)
        # geometric(1/2) length: 1 + number of trailing one bits, capped
        length = np.ones(n, dtype=np.int64)
        bits = r3.copy()
        alive = np.ones(n, dtype=bool)
        for _ in range(_MAX_INS - 1):
            alive &= (bits & np.uint64(1)) == 1
            if not alive.any():
                break
            length += alive
            bits >>= np.uint64(1)
        sub_shift = 1 + (_stream(seed, 4, n) % np.uint64(3)).astype(np.int64)
    is_del = ev & (kind == 9)
    # positions covered by any deletion
    diff = np.zeros(n + _MAX_INS + 2, dtype=np.int64)
    starts = np.nonzero(is_del)[0]
    np.add.at(diff, starts, 1)
    np.add.at(diff, starts + length[starts], -1)
    covered = np.cumsum(diff)[:n] > 0
    is_sub = ev & (kind < 8) & ~covered
    is_ins = ev & (kind == 8) & ~covered
    qcode = np.where(is_sub, (tcode + sub_shift) % 4, tcode)
    ins_len = np.where(is_ins, length, 0)
    keep = (~covered).astype(np.int64)
    out_cnt = ins_len + keep
    offs = np.concatenate(([0], np.cumsum(out_cnt)))
    ql = int(offs[-1])
    q = np.empty(ql, dtype=np.uint8)
    # kept target positions go at offs[i] + ins_len[i]
    kp = np.nonzero(keep)[0]
    q[offs[kp] + ins_len[kp]] = _BASES[qcode[kp]]
    # inserted bases
    ip = np.nonzero(is_ins)[0]
    if ip.size:
        rep = np.repeat(ip, ins_len[ip])
        within = np.arange(rep.size) - np.repeat(np.cumsum(ins_len[ip]) - ins_len[ip], ins_len[ip])
        with np.errstate(over="ignore"):
            key = _splitmix64(np.array([(seed * 0x100 + 5) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
            rb = _splitmix64(key + (rep.astype(np.uint64) * np.uint64(_MAX_INS) + within.astype(np.uint64)) * np.uint64(0xD1342543DE82EF95))
        q[offs[rep] + within] = _BASES[(rb >> np.uint64(62)).astype(np.int64)]
    return _BASES[tcode].tobytes(), q.tobytes()


def make_batch(n_pairs, n, p, first_index=0):
    return [make_pair(n, p, first_index + i) for i in range(n_pairs)]
