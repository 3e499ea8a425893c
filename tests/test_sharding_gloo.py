"""world_size-2 gloo test of the data-parallel host logic (no GPU): contiguous shard ranges and
the rank-ordered gather of ragged transcripts.  The per-rank 'model' is a stub that tags its
rank, so only the sharding/gather plumbing is under test."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StubModel:
    def transcribe(self, audio):
        # token list derived from the data only: (first sample * 1000, length marker ...)
        return [[int(round(float(a[0]) * 1000))] * (1 + int(round(float(a[0]) * 1000)) % 3) for a in audio]


def _worker(rank, world, port, n_items, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "velocity-asr_b200"))
    from velocity_asr.sharding import shard_range, transcribe_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    audio = torch.arange(n_items, dtype=torch.float32).unsqueeze(1).repeat(1, 4) / 1000.0
    res = transcribe_sharded(_StubModel(), audio)
    # the compact exchange: fixed-size token matrices to rank 0
    from velocity_asr.sharding import gather_token_arrays
    import numpy as np
    tok = np.full((3, 5), rank + 1, dtype=np.int32)
    cnt = np.array([rank, 2, 5], dtype=np.int32)
    arr = gather_token_arrays(tok, cnt, dst=0)
    if rank == 0:
        assert arr[0].shape == (3 * world, 5) and arr[1].tolist() == [0, 2, 5, 1, 2, 5][:3 * world]
        assert (arr[0][3:] == 2).all() and (arr[0][:3] == 1).all()
    else:
        assert arr is None
    out[rank] = (shard_range(n_items, rank, world), res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 2, 1])
def test_two_rank_shard_and_gather(n_items):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, out), nprocs=world, join=True)
    audio = torch.arange(n_items, dtype=torch.float32).unsqueeze(1).repeat(1, 4) / 1000.0
    expect = _StubModel().transcribe(audio)
    ranges = [out[r][0] for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_items and ranges[0][1] == ranges[1][0]
    for r in range(world):
        assert out[r][1] == expect


def test_shard_range_covers_everything():
    import sys
    from velocity_asr.sharding import shard_range
    for n in (0, 1, 5, 64, 511, 512):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
