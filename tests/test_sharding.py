"""N > 1 host logic on CPU (gloo, world_size 2 and 3): line-aligned sharding, histogram all-reduce, global row bases.
The per-shard engine here is the oracle (no GPU in this container); on the GPU box the same functions run over NCCL in
bench.py, and tests/test_gpu_parity.py::test_multi_device_engine covers the in-library multi-device path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_line_aligned_ranges_partition_the_text():
    from gorp_b200 import corpus, sharding
    text = corpus.readme_corpus(5000, seed=3)
    for world in (1, 2, 3, 8):
        cuts = [sharding.line_aligned_range(text, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == text.size
        for (a, b), (c, d) in zip(cuts, cuts[1:]):
            assert b == c
        for a, b in cuts:
            assert a == 0 or text[a - 1] == 0x0A
    # text without a trailing '\n', a single huge line, empty text
    t2 = text[:-1]
    assert sharding.line_aligned_range(t2, 1, 2)[1] == t2.size
    one = np.full(1000, ord("a"), dtype=np.uint16)
    assert [sharding.line_aligned_range(one, r, 4) for r in range(4)] == [(0, 1000), (1000, 1000), (1000, 1000), (1000, 1000)]
    assert sharding.line_aligned_range(np.zeros(0, np.uint16), 0, 2) == (0, 0)
    assert [sharding.line_range(10, r, 3) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gorp_b200 import corpus, sharding
    from oracle import gorp_oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        d, gen = corpus.CONFIGS["readme"]
        text = gen(20000, seed=11)[:-1]  # the last line is not terminated
        o = gorp_oracle.Gorp(d)
        u0, u1 = sharding.line_aligned_range(text, rank, world)
        shard = text[u0:u1]
        st, en = gorp_oracle.split_lines(shard)
        oe, osp = o.extract_batch(shard, (st, en), threads=1)
        E = len(o.extractions)
        hist = torch.zeros(E + 2, dtype=torch.int64)
        for e in oe.tolist():
            hist[e if e >= 0 else (E if e == -1 else E + 1)] += 1
        sharding.allreduce_histogram(hist)
        base, total = sharding.global_row_base(len(st))
        q.put((rank, base, total, hist.tolist(), sharding.rebase_line_offsets(st, u0), oe, osp))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_extraction_matches_the_whole_batch(world):
    import torch.multiprocessing as mp
    from gorp_b200 import corpus
    from oracle import gorp_oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, gen = corpus.CONFIGS["readme"]
    text = gen(20000, seed=11)[:-1]
    o = gorp_oracle.Gorp(d)
    st, en = gorp_oracle.split_lines(text)
    oe, osp = o.extract_batch(text, (st, en), threads=1)
    E = len(o.extractions)
    want_hist = np.bincount(np.where(oe >= 0, oe, np.where(oe == -1, E, E + 1)), minlength=E + 2)
    rows = 0
    for rank, base, total, hist, starts, e, sp in got:
        assert base == rows and total == len(st) and hist == want_hist.tolist()
        n = len(starts)
        assert (starts == st[rows:rows + n]).all() and (e == oe[rows:rows + n]).all() and (sp == osp[rows:rows + n]).all()
        rows += n
    assert rows == len(st)
