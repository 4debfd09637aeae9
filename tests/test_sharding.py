"""world_size-2 gloo tests (CPU) of the batch-sharding host logic: each rank runs its
shard, outputs are all-gathered for validation only and must equal the un-sharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megengine_b200 import _capi as C
from megengine_b200 import sharding as S
from . import _cases as K
from . import _oracle as O


def test_shard_range_covers_batch_exactly():
    for n in (0, 1, 7, 8, 255, 256, 8192):
        for world in (1, 2, 3, 4, 8):
            spans = [S.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _c1) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        S.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d_full = K.make_desc("qs8", n_total, 16, 6, 6, 16, 3, pad=1, nonline=C.NL_RELU, has_z=True)
        src, flt, bias, z = K.gen_inputs(d_full, 31)
        tsrc = torch.from_numpy(src).reshape(n_total, 6, 6, 16)
        tz = torch.from_numpy(z).reshape(n_total, 6, 6, 16)
        lsrc, lz = S.shard_batch(tsrc, world, rank), S.shard_batch(tz, world, rank)
        start, cnt = S.shard_range(n_total, world, rank)
        assert lsrc.shape[0] == cnt
        d = d_full.copy()
        d.n = cnt
        # the per-rank "device work" is played by the oracle here (CPU test of the host logic only)
        ldst = O.oracle_convbias(d, lsrc.contiguous().numpy().reshape(-1), flt, bias, lz.contiguous().numpy().reshape(-1))
        ldst = torch.from_numpy(ldst).reshape(cnt, 6, 6, 16)
        full = S.gather_batch(ldst, n_total)
        sums = S.gather_checksums(ldst)
        want = O.oracle_convbias(d_full, src, flt, bias, z).reshape(n_total, 6, 6, 16)
        ok = np.array_equal(full.numpy(), want)
        want_sums = [S.checksum(S.shard_batch(torch.from_numpy(want), world, r)) for r in range(world)]
        q.put((rank, ok, sums == want_sums))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [4, 5])
def test_two_rank_sharded_convbias_matches_unsharded(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(ok and sums_ok for _r, ok, sums_ok in res), res
