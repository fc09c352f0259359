"""Frame sharding of the MSDeformAttn core across ranks (one process per GPU).

The op has no cross-frame dependency (ops/src/cuda/ms_deform_im2col_cuda.cuh:268-274: the batch index only selects the value
slab), so a batch shards by clip exactly as the reference's DDP data loader does it: SOLVER.IMS_PER_BATCH clips are divided by
the world size (train_net.py:117) and each clip contributes NUM_FRAMES frames (maskformer_model.py:316-323).  There is no
collective on the data path; the only cross-rank traffic of the bench is the max-reduction of the timings."""
import torch
import torch.distributed as dist


def shard_clips(n_clips, frames_per_clip, rank, world):
    """-> (first_frame, n_frames) of the contiguous block of whole clips owned by `rank`.  Clips are dealt out in contiguous
    blocks; when n_clips is not divisible by world the first n_clips % world ranks get one extra clip."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_clips, world)
    mine = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first * frames_per_clip, mine * frames_per_clip


def max_over_ranks(value, device=None):
    """Max-reduce a python float over the default process group (identity when not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
