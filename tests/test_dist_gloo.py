"""world_size-2 gloo tests (CPU) of the multi-rank algorithm of the loss path.

The CUDA path shards images over ranks and exchanges only (i) the gradient arena, (ii) the [K] center batch sums and
(iii) the [K] Sinkhorn prototype sums.  These tests run the *algorithm* the kernels implement (log-domain diagonal
scaling, lazy center EMA) with real gloo all-reduces and check it against the reference formulation
(oracle.sinkhorn_knopp / oracle.center_ema) evaluated on the concatenated global batch in one process.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dinov2_oracle as O


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def logdomain_sinkhorn(t: torch.Tensor, t_scale: float, n_iter: int = 3):
    """Pure-torch statement of lightly_train_b200 sinkhorn_colterm + row_lse (what the kernels compute)."""
    z = t.float() * t_scale
    logv = torch.zeros(t.shape[0])
    logu = None
    for it in range(n_iter):
        sums = torch.exp(z + logv[:, None]).sum(0)
        if dist.is_initialized():
            dist.all_reduce(sums)
        logu = -torch.log(sums)
        if it + 1 < n_iter:
            logv = -torch.logsumexp(z + logu[None, :], dim=1)
    rowterm = -torch.logsumexp(z + logu[None, :], dim=1)
    return torch.exp(z + logu[None, :] + rowterm[:, None])


def _worker(rank: int, world: int, port: int, tmp: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(77)
    K, rows = 64, 10
    t_all = torch.randn(world * rows, K, generator=g)
    mine = t_all[rank * rows:(rank + 1) * rows]
    # Sinkhorn: sharded log-domain iteration == reference formulation on the global batch
    q_mine = logdomain_sinkhorn(mine, 1 / 0.05)
    q_ref = O.sinkhorn_knopp(t_all, 0.05, float(world * rows))[rank * rows:(rank + 1) * rows]
    torch.testing.assert_close(q_mine, q_ref, rtol=1e-4, atol=1e-7)
    # the reference's own multi-rank form (all_reduce hooks) agrees as well
    q_ref2 = O.sinkhorn_knopp(mine, 0.05, float(world * rows), all_reduce=lambda x: (dist.all_reduce(x), x)[1])
    torch.testing.assert_close(q_mine, q_ref2, rtol=1e-4, atol=1e-7)
    # center EMA: per-rank sums, all-reduced, divided by len*world
    center = torch.zeros(1, K)
    s = O.center_batch_sum_dino(mine)
    dist.all_reduce(s)
    c_dist = O.center_ema(center, s, rows * world, 0.9)
    c_ref = O.center_ema(center, O.center_batch_sum_dino(t_all), rows * world, 0.9)
    torch.testing.assert_close(c_dist, c_ref, rtol=1e-5, atol=1e-7)
    # gradient arena all-reduce(sum) with grad_scale = 1/world == DDP mean
    grad = torch.full((1000,), float(rank + 1))
    dist.all_reduce(grad)
    assert torch.allclose(grad / world, torch.full((1000,), 1.5))
    dist.destroy_process_group()
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("ok")


def test_world2_loss_path_algorithm(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_logdomain_sinkhorn_single_process_matches_reference_form():
    t = torch.randn(12, 32)
    torch.testing.assert_close(logdomain_sinkhorn(t, 20.0), O.sinkhorn_knopp(t, 0.05, 12.0), rtol=1e-4, atol=1e-7)
