"""Multi-GPU sharding of a neutron batch: every neutron is independent, tables are
replicated per GPU, the GLOBAL neutron index range is partitioned contiguously over the ranks and
the random stream of a neutron is keyed by its global index -- so the union of the shards is
bit-identical to a single-GPU run for any world size.  The only collective is the element-wise sum
of tally arrays (the analogue of the reference's Tally::merge, NCMMC_Tally.hh:40-62)."""


def shard_range(n_total, rank, world):
    """Contiguous [begin, end) of rank `rank`; sizes differ by at most one, order preserved."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_tallies(hist, group=None):
    """In-place sum of a tally tensor over all ranks (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist
