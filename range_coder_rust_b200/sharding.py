"""Multi-GPU plumbing: chunks shard across ranks with no data-path collective; the
only exchange on the path is the all-reduce of the K-entry symbol-count table when a
static global model is used.  torch.distributed is plumbing here: one
process per GPU, NCCL on GPUs (gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_chunks(n_chunks, rank, world):
    """Contiguous block partition: rank r owns chunks [lo, hi)."""
    lo = (n_chunks * rank) // world
    hi = (n_chunks * (rank + 1)) // world
    return lo, hi


def shard_symbols(n_syms, chunk_syms, rank, world):
    """Symbol range of rank r's chunks (the last shard absorbs a ragged final chunk)."""
    n_chunks = (n_syms + chunk_syms - 1) // chunk_syms
    lo, hi = shard_chunks(n_chunks, rank, world)
    return lo * chunk_syms, min(hi * chunk_syms, n_syms)


def init_comm(ctx):
    """The library's own NCCL communicator for this process group (include/rcb200.h, rcb_comm_*):
    torch.distributed only ships rank 0's 128-byte unique id to the other ranks; the all-reduce on
    the path is then issued by librcb200.so itself (`Context.allreduce_counts`).  Returns None for a
    single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    from .api import Comm

    world, rank = dist.get_world_size(), dist.get_rank()
    box = [Comm.unique_id(ctx.lib) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return ctx.comm_init_rank(box[0], world, rank)


def allreduce_counts(counts, ctx=None, comm=None):
    """Sum the local u64 histograms (stored as int64 bit patterns) over all ranks, in place.
    With a library communicator (`init_comm`) this is rcb_allreduce_counts -- NCCL driven through the
    C ABI, on the context's stream; without one (the gloo tests of the host logic on CPU) it falls to
    torch.distributed.  Counts never exceed 2^63 in practice (8 EiB of symbols), so gloo's signed sum
    is exact too."""
    assert counts.dtype == torch.int64
    if comm is not None:
        ctx.allreduce_counts(counts, comm)
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def global_offsets(local_stream_bytes):
    """Exclusive prefix sum of the per-rank compressed sizes: where each rank's segment
    starts in the concatenated stream (an all-gather of one int64 per rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, int(local_stream_bytes)
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(local_stream_bytes)], dtype=torch.int64)
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        mine = mine.to(dev)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [int(s.item()) for s in sizes]
    return sum(sizes[:rank]), sum(sizes)
