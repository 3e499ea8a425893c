"""Data-parallel sharding of utterance batches: one process per GPU, no collective on the data
path (every utterance is independent end to end: per-utterance mel statistics, audio.py:132-135,
and no cross-batch op in model.py / ssm.py / attention.py), only a host-side gather of the
ragged transcripts in rank order."""
from typing import List, Sequence, Tuple


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split: rank r owns [start, end); the first n % world ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_transcripts(local: Sequence[List[int]], group=None) -> List[List[int]]:
    """All ranks receive the token lists of the whole batch, in batch order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [list(t) for t in local]
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, [list(t) for t in local], group=group)
    return [t for part in parts for t in part]


def gather_token_arrays(tokens, lens, group=None, dst: int = 0):
    """The same exchange in compact form, for a serving loop: every rank passes its (B, L) left-packed int32 token
    matrix and (B,) counts (host numpy arrays or CPU tensors, e.g. from transcribe_batches(..., as_arrays=True));
    rank `dst` receives the (world * B, L) / (world * B,) stack in rank order, the others None.  Fixed-size
    tensors over the host (gloo) group: no pickling, no Python ints (all_gather_object of 8 x 64 token lists costs
    milliseconds per step in unpickling alone).  B and L must be equal on all ranks."""
    import torch
    import torch.distributed as dist
    tok = torch.as_tensor(tokens).to(torch.int32).contiguous()
    cnt = torch.as_tensor(lens).to(torch.int32).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tok.numpy(), cnt.numpy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    root = dist.get_global_rank(group, dst) if group is not None else dst
    if rank == dst:
        all_tok = torch.empty((world,) + tuple(tok.shape), dtype=torch.int32)
        all_cnt = torch.empty((world,) + tuple(cnt.shape), dtype=torch.int32)
        dist.gather(tok, list(all_tok.unbind(0)), dst=root, group=group)
        dist.gather(cnt, list(all_cnt.unbind(0)), dst=root, group=group)
        return all_tok.reshape(-1, tok.shape[-1]).numpy(), all_cnt.reshape(-1).numpy()
    dist.gather(tok, None, dst=root, group=group)
    dist.gather(cnt, None, dst=root, group=group)
    return None


def transcribe_sharded(model, audio, group=None) -> List[List[int]]:
    """audio (B, S) host tensor, identical on every rank -> transcripts of all B utterances.
    Each rank runs model.transcribe on its own slice (model already on this rank's GPU)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    s, e = shard_range(audio.shape[0], rank, world)
    local = model.transcribe(audio[s:e]) if e > s else []
    return gather_transcripts(local, group)
