"""CPU tests of the N>1 host path (adamvs_b200/sharding.py): partition arithmetic and a world_size-2
gloo run of the shard -> per-view work -> host gather loop (the path shards whole reference views; there
is no data-path collective to test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from adamvs_b200 import sharding


@pytest.mark.parametrize("n,world", [(256, 8), (7, 2), (3, 4), (0, 2), (1, 1), (10, 3)])
def test_shard_range_is_a_balanced_partition(n, world):
    spans = [sharding.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (b0, e0), (b1, e1) in zip(spans, spans[1:]):
        assert e0 == b1
    sizes = [e - b for b, e in spans]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_batches_cover_slice_with_ragged_tail():
    assert sharding.batches(3, 14, 4) == [(3, 7), (7, 11), (11, 14)]
    assert sharding.batches(5, 5, 8) == []
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def run_batch(b, e):                      # stand-in for one forward over views [b, e)
            return [torch.full((2, 3), float(i)) + rank * 0.0 for i in range(b, e)]
        res = sharding.run_sharded(n_units, 2, run_batch)
        worst = sharding.max_over_ranks(10.0 + rank)
        if rank == 0:
            torch.save({"res": res, "worst": worst}, out_path)
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [7, 2, 1])
def test_two_rank_gloo_gather_preserves_view_order(tmp_path, n_units):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), n_units, out), nprocs=2, join=True)
    got = torch.load(out)
    assert got["worst"] == 11.0
    assert len(got["res"]) == n_units
    for i, t in enumerate(got["res"]):
        assert torch.equal(t, torch.full((2, 3), float(i)))
