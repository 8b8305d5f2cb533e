"""world_size-2 gloo test of the data-parallel plumbing (flat bucket, sharding, sum-then-identical
update).  CPU only: the kernels are not exercised, only the exchange step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mgr_b200 import parallel
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_rows(10, rank, world)
    params = [torch.zeros(3, 4), torch.zeros(5)]
    bucket = parallel.FlatGradBucket(params)
    g = torch.Generator().manual_seed(0)
    full = [torch.randn(10, 3, 4, generator=g), torch.randn(10, 5, generator=g)]
    local = [f[lo:hi].sum(0) for f in full]          # per-rank partial gradient of a sum over rows
    bucket.pack(local)
    views = bucket.all_reduce()
    ok = all(torch.allclose(v, f.sum(0), atol=1e-6) for v, f in zip(views, full))
    out.put((rank, lo, hi, ok))
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert [r[1:3] for r in res] == [(0, 5), (5, 10)]
    assert all(r[3] for r in res)


def test_shard_rows_cover_batch():
    from mgr_b200 import parallel
    for n, w in [(256, 8), (256, 3), (7, 4)]:
        spans = [parallel.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
