"""CPU, world_size 2, gloo: the N>1 path of the sampler driver -- batch sharding with no data-path
collective and one all-gather of finished samples at the end -- gives the same result as one rank."""
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


def _fake_chain(cond, lo, hi):
    """Stand-in for a per-sample sampling chain: depends only on the global sample index and cond."""
    from polyffusion_b200.parallel import sample_seed

    out = []
    for i, gi in enumerate(range(lo, hi)):
        g = torch.Generator().manual_seed(sample_seed(3, gi))
        out.append(torch.randn(2, 8, 8, generator=g) + cond[i].sum())
    return torch.stack(out) if out else torch.zeros(0, 2, 8, 8)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from polyffusion_b200.parallel import sample_sharded

    cond = torch.arange(total * 3, dtype=torch.float32).reshape(total, 1, 3)
    full = sample_sharded(_fake_chain, cond, total)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def _run(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cond = torch.arange(total * 3, dtype=torch.float32).reshape(total, 1, 3)
    want = _fake_chain(cond, 0, total)
    for r in range(2):
        assert res[r].shape == want.shape and torch.equal(res[r], want)


def test_two_rank_gather_even():
    _run(8)


def test_two_rank_gather_ragged():
    _run(5)
