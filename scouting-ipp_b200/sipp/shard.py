"""Multi-GPU plumbing of the evaluation step: candidates are sharded in contiguous blocks, the map is
replicated, and the only cross-rank traffic is one all_gather of a 16-byte (value: f64, index: i64) record per rank.
torch.distributed is used as plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous block of candidates owned by `rank`: sizes differ by at most one, earlier ranks get the longer blocks."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_best(value, index, device=None):
    """16-byte record as an int64[2] tensor: [f64 bit pattern of value, global index]."""
    rec = torch.empty(2, dtype=torch.int64, device=device)
    rec[0:1].view(torch.float64)[0] = float(value)
    rec[1] = int(index)
    return rec


def gather_best(local_rec, world, group=None, out=None):
    """all_gather the per-rank records and pick the winner exactly like a sequential scan over the concatenated candidate
    list would: strictly greater value wins, ties keep the lowest global index, NaN never beats a number, a rank with no
    candidate (index < 0) never wins. Returns (value, global_index) as a 2-element int64 record tensor."""
    if world == 1:
        return local_rec
    if out is None:
        out = torch.empty(2 * world, dtype=torch.int64, device=local_rec.device)
    dist.all_gather_into_tensor(out, local_rec, group=group)
    recs = out.view(world, 2)
    vals = recs[:, 0].view(torch.float64)
    idx = recs[:, 1]
    valid = (idx >= 0) & ~torch.isnan(vals)
    vals = torch.where(valid, vals, torch.full_like(vals, float("-inf")))
    best = vals.max()
    # lowest global index among the ranks holding the maximum; with no comparable value anywhere every entry ties at
    # int64 max and rank 0's record is returned (the sequential scan keeps its first element). No host synchronisation.
    cand = torch.where(valid & (vals == best), idx, torch.full_like(idx, torch.iinfo(torch.int64).max))
    j = torch.argmin(cand)
    return recs.index_select(0, j.view(1)).view(2)


def ensemble_shard(items, rank, world):
    """Maps of an ensemble owned by `rank`: every world-th one starting at `rank` (maps are independent — no data-path
    communication; only the timing is reduced over the ranks)."""
    return list(items)[int(rank)::int(world)]


def job_throughput(units_per_rank, seconds_local, world, group=None):
    """Whole-job rate of a sharded run: all ranks' units over the slowest rank's time (all_reduce MAX of the local time)."""
    t = torch.tensor([float(seconds_local)], dtype=torch.float64)
    if world > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(units_per_rank) * world / float(t.item()), float(t.item())
