"""Multi-GPU partitioning of the commitment path.  One process per GPU; the path shards
with no data-path collective except the combine of per-rank partial commitments of ONE large MSM:

    large MSM      contiguous point ranges  -> all_gather of one 64-byte affine point per rank -> local sum
    batches        contiguous batch ranges  -> nothing to exchange (optionally all_gather the results)

Group addition is not a reduction operator NCCL knows, hence all_gather + a local point-add kernel
(vkzg_g1_sum_dev) instead of all_reduce.  The helpers take the torch.distributed module so that the same code
runs over NCCL on GPUs and over gloo in the CPU tests.
"""


def split_range(total, world, rank):
    """contiguous [first, first + count) of `total` units for `rank`; the first total % world ranks get one more"""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_gather_points(dist, torch, part):
    """part: uint8 [64] / [k, 64] tensor (this rank's partial commitments) -> uint8 [world, k, 64] on every rank"""
    world = dist.get_world_size()
    part = part.reshape(-1, 64).contiguous()
    out = torch.empty((world,) + tuple(part.shape), dtype=torch.uint8, device=part.device)
    dist.all_gather_into_tensor(out.view(world * part.shape[0], 64), part)
    return out


def combine_partial_msm(dist, torch, engine, part):
    """all ranks end with the full MSM result: gather the per-rank partial points, add them on the device"""
    allp = all_gather_points(dist, torch, part)
    out = torch.empty((1, 64), dtype=torch.uint8, device=part.device)
    engine.g1_sum_dev(allp.view(-1, 64), allp.shape[0] * allp.shape[1], out)
    return out
