"""world_size-2 `gloo` test of the N>1 path on CPU: sharding + the single all-gather of final latents give the
same tensor as the single-process run (the per-rank sampler is a deterministic stand-in; the CUDA engine is
covered by the -m gpu tests)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _stand_in_sampler(z, y):
    # deterministic, per-sample function of (z_i, y_i) only -- like the real sampler, no cross-sample coupling
    return torch.tanh(z * 1.5) + y.view(-1, 1, 1, 1).float() * 0.01


def _worker(rank, world, port, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from va_vae_b200.dist import sample_sharded
        g = torch.Generator().manual_seed(1234)
        z = torch.randn(n, 4, 4, 4, generator=g)
        y = torch.randint(0, 31, (n,), generator=torch.Generator().manual_seed(4321))
        out = sample_sharded(_stand_in_sampler, z, y)
        if rank == 0:
            torch.save(out, out_path)
        # every rank must hold the full gathered result
        ref = _stand_in_sampler(z, y)
        assert torch.equal(out, ref)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7, 1])      # even, ragged, and n < world size (rank 1 holds an empty shard)
def test_sharded_sampling_world2_equals_single_process(tmp_path, n):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, port, n, out_path), nprocs=2, join=True)
    out = torch.load(out_path)
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(n, 4, 4, 4, generator=g)
    y = torch.randint(0, 31, (n,), generator=torch.Generator().manual_seed(4321))
    assert torch.equal(out, _stand_in_sampler(z, y))


def test_shard_ranges_cover_every_sample_exactly_once():
    from va_vae_b200.dist import shard_range
    for n in (0, 1, 7, 128, 1024, 1031):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
