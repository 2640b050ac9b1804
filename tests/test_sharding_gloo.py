"""CPU: the N>1 host path (slice sharding + result gather) on a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miccai24_immoco_b200.sharding import gather_images, reconstruct_slices, shard_indices


def test_shard_indices_partition():
    for n in (0, 1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            got = sorted(i for r in range(world) for i in shard_indices(n, r, world))
            assert got == list(range(n))
            sizes = [len(shard_indices(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_reconstruct(kspace, masks, iters, lr, lam, debug):
    # stands in for the CUDA fit: a deterministic function of the inputs, so order mix-ups show
    return kspace * (1 + masks.shape[0]) + iters, kspace


def _worker(rank, world, port, n_slices, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        ks = [torch.complex(torch.randn(6, 4, generator=g), torch.randn(6, 4, generator=g)) for _ in range(n_slices)]
        ms = [torch.zeros((1 + s % 3, 6, 4), dtype=torch.long) for s in range(n_slices)]
        out = reconstruct_slices(ks, ms, iters=7, reconstruct_fn=_fake_reconstruct)
        if rank == 0:
            want = torch.stack([_fake_reconstruct(ks[s], ms[s], 7, 0, 0, False)[0] for s in range(n_slices)])
            ret["ok"] = bool(out is not None and out.shape == want.shape and torch.allclose(out, want))
        else:
            ret[f"none{rank}"] = out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_slices", [5, 4, 1])
def test_reconstruct_slices_world2_gloo(n_slices):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_slices, ret), nprocs=2, join=True)
    assert ret.get("ok") is True
    assert ret.get("none1") is True


def test_gather_single_rank_is_identity():
    x = torch.complex(torch.randn(3, 4, 4), torch.randn(3, 4, 4))
    assert gather_images(x, 3, 0, 1) is x
