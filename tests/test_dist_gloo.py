"""CPU, world_size 2 over gloo: the N>1 host path (shard -> align locally -> one gather to rank 0).  The local aligner
is injected: here the oracle stands in for the GPU engine, so only sharding, packing and the collective are under test."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_align(opt, pairs):
    from oracle import orc
    return [orc.oracle_exact(orc.copy_opt(opt), t, q) for t, q in pairs]


def _worker(rank, world, port, balance, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from miniwfa_b200 import dist as mdist, synth
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs = [synth.make_pair(n, 0.08, 500 + i) for i, n in enumerate([0, 40, 900, 1, 300, 2000, 17, 650, 5])]
        pairs[0] = (b"ACGT", b"")  # empty query
        for kw in ({}, {"flag": 1}, {"flag": 1, "step": 9}, {"flag": 1, "max_s": 60}):
            opt = orc.make_opt(**kw)
            got = mdist.wfa_exact_batch_sharded(opt, pairs, align_fn=_oracle_align, balance=balance)
            if rank == 0:
                want = _oracle_align(opt, pairs)
                assert got == want, kw
            else:
                assert got is None
        # score-only batches: the results travel in ONE collective (a gather of equal-sized records)
        idx = mdist.shard_indices(len(pairs), world, rank)
        local = _oracle_align(orc.make_opt(), [pairs[i] for i in idx])
        got = mdist.gather_to_root(idx, local, len(pairs), fixed=True)
        assert (got == _oracle_align(orc.make_opt(), pairs)) if rank == 0 else got is None
        # an empty batch and a batch smaller than the world
        assert mdist.wfa_exact_batch_sharded(orc.make_opt(), [], align_fn=_oracle_align) in ([], None)
        one = mdist.wfa_exact_batch_sharded(orc.make_opt(flag=1), pairs[2:3], align_fn=_oracle_align)
        if rank == 0:
            assert one == _oracle_align(orc.make_opt(flag=1), pairs[2:3])
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL %r" % (e,)))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balance", [False, True])
def test_sharded_batch_world2_gloo(balance):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, balance, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, "ok"), (1, "ok")], res
    assert all(p.exitcode == 0 for p in procs)


def test_shard_indices_partition():
    from miniwfa_b200.dist import shard_indices, pack_results, unpack_results
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 128, 1024):
            seen = sorted(i for r in range(world) for i in shard_indices(n, world, r))
            assert seen == list(range(n))
            costs = [(i * 7919) % 101 + 1 for i in range(n)]
            parts = [shard_indices(n, world, r, costs) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            if n >= 8 * world:
                loads = [sum(costs[i] for i in p) for p in parts]
                assert max(loads) - min(loads) <= max(costs)
    res = [(5, 2, 99, [0x17, 0x28]), (-1, 0, 12, []), (0, 0, 0, [])]
    out = unpack_results(pack_results([4, 0, 2], res), [None] * 5)
    assert out == [res[1], None, res[2], None, res[0]]
