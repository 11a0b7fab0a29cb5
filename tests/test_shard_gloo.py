"""N > 1 path on CPU: two gloo ranks deal a batch of samples (samples shard with no data-path
collective), each computes its share (here with the CPU oracle standing in for the GPU call — test infrastructure
only), and rank 0 reassembles every result in sample order."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from dipgenie_b200 import shard, synth


def test_shard_indices_partition():
    for n in (0, 1, 5, 22, 23):
        for world in (1, 2, 3, 8):
            seen = sorted(i for r in range(world) for i in shard.shard_indices(n, world, r))
            assert seen == list(range(n))
            sizes = [len(shard.shard_indices(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)
    with pytest.raises(ValueError):
        shard.merge_shards(3, [[1], [2]], 2)


def _samples():
    out = []
    for seed in range(7):
        rng = np.random.default_rng(300 + seed)
        g = synth.random_level_graph(40 + seed, n_levels=int(rng.integers(3, 12)), max_width=int(rng.integers(2, 7)),
                                     n_colours=int(rng.integers(0, 40)), p_colour=0.5)
        out.append((g, int(rng.integers(0, 4))))
    return out


def _oracle_local(samples):
    import oracle
    res = []
    for g, R in samples:
        o = oracle.dp_diploid(g.level_off, g.adj_off, g.adj_dst, g.adj_w, g.col_off, g.col_val, g.colour_is_hom, R,
                              want_checksums=False)
        res.append((int(o["value"]), int(o["s_het"]), o["p1_edges"].tolist(), o["p2_edges"].tolist()))
    return res


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = shard.run_sharded(_samples(), _oracle_local, dist)
        if rank == 0:
            q.put(res)
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_reassemble_batch():
    import oracle
    oracle.build()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got == _oracle_local(_samples())
    assert shard.run_sharded(_samples(), _oracle_local, None) == got      # no process group: plain local call
