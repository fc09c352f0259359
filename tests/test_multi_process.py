"""CPU, world_size 2 over gloo: the N>1 path of the bench -- sharding frames by clip with no data-path collective, and the
max-over-ranks reduction of timings.  The per-rank compute is the CPU oracle (the CUDA kernels need a GPU); what is tested
is that shards are disjoint, cover the batch, and that per-rank results concatenate to the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import S4_SHAPES, make_inputs


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_clips, frames, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from combo_avs_b200.sharding import max_over_ranks, shard_clips
        d = make_inputs(n_clips * frames, S4_SHAPES, 2, 32, 40, 4, seed=17, lo=-0.1, hi=1.1, dtype=np.float64)
        first, count = shard_clips(n_clips, frames, rank, world)
        sl = slice(first, first + count)
        out = oracle.forward(d["value"][sl], d["shapes"], d["start"], d["loc"][sl], d["attn"][sl])
        gv, gl, ga = oracle.backward(d["value"][sl], d["shapes"], d["start"], d["loc"][sl], d["attn"][sl], d["grad_out"][sl])
        np.savez(os.path.join(tmpdir, "rank%d.npz" % rank), first=first, count=count, out=out, gv=gv, gl=gl, ga=ga)
        # timing reduction: every rank must see the slowest rank's time
        slowest = max_over_ranks(10.0 + rank)
        assert slowest == 10.0 + (world - 1)
        # no collective on the data path: the only tensors that crossed ranks are the 8-byte timing scalars
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_two_ranks_shard_by_clip(tmp_path, n_clips):
    import __graft_entry__ as g
    g.build()
    world, frames = 2, 5
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_clips, frames, str(tmp_path)), nprocs=world, join=True)
    import oracle
    d = make_inputs(n_clips * frames, S4_SHAPES, 2, 32, 40, 4, seed=17, lo=-0.1, hi=1.1, dtype=np.float64)
    full_out = oracle.forward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"])
    full_gv, full_gl, full_ga = oracle.backward(d["value"], d["shapes"], d["start"], d["loc"], d["attn"], d["grad_out"])
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    covered = np.zeros(n_clips * frames, dtype=int)
    for p in parts:
        f, c = int(p["first"]), int(p["count"])
        assert f % frames == 0 and c % frames == 0  # whole clips only
        covered[f:f + c] += 1
        assert np.array_equal(p["out"], full_out[f:f + c])
        assert np.array_equal(p["gv"], full_gv[f:f + c])
        assert np.array_equal(p["gl"], full_gl[f:f + c])
        assert np.array_equal(p["ga"], full_ga[f:f + c])
    assert (covered == 1).all()


def test_shard_clips_arithmetic():
    from combo_avs_b200.sharding import shard_clips
    for n_clips, world in [(8, 1), (8, 2), (8, 4), (8, 8), (5, 2), (3, 4), (0, 2)]:
        seen = []
        for r in range(world):
            first, count = shard_clips(n_clips, 5, r, world)
            seen.extend(range(first, first + count))
        assert seen == list(range(n_clips * 5))
    # global batch 8 clips x 5 frames (S4): 40 / 20 / 10 / 5 frames per rank at 1 / 2 / 4 / 8 GPUs
    assert [shard_clips(8, 5, 0, w)[1] for w in (1, 2, 4, 8)] == [40, 20, 10, 5]
    with pytest.raises(ValueError):
        shard_clips(8, 5, 2, 2)
