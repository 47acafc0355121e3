"""Sharding helpers for multi-GPU runs (one process per GPU).

The hot path has no cross-sample state, so shards are independent and need no collective.  Partitioning follows
the reference's reader convention (dali/operators/reader/loader/loader.cc:78-81, plugin/base_iterator.py:305-310): shard s of S
over N samples covers [N*s/S, N*(s+1)/S).  `all_gather_output` is the optional step for consumers that need the full batch on
every GPU (BASELINE config 5; the reference has no call site for it): one in-place NCCL all-gather of the fp16 NCHW tensor.
"""


def shard_range(num_samples, shard_id, num_shards):
    if not (0 <= shard_id < num_shards):
        raise ValueError(f"shard_id {shard_id} out of range for {num_shards} shards")
    return num_samples * shard_id // num_shards, num_samples * (shard_id + 1) // num_shards


def shard_of(items, shard_id, num_shards):
    lo, hi = shard_range(len(items), shard_id, num_shards)
    return items[lo:hi]


def sharded_source(items, batch_size, shard_id, num_shards):
    """An fn.external_source callable returning this rank's slice, one batch per iteration (cycling)."""
    mine = shard_of(items, shard_id, num_shards)
    if not mine:
        raise ValueError("empty shard")

    def source(iteration):
        start = (iteration * batch_size) % len(mine)
        return [mine[(start + k) % len(mine)] for k in range(batch_size)]
    return source


class GatherBuffer:
    """One persistent [world * B_local, ...] tensor per rank.  `local` is this rank's slice of it: bind it as the output of the
    last operator (ImagePipelineC2.bind_output) so that CropMirrorNormalize writes its batch where the collective expects it;
    `all_gather()` is then ONE in-place NCCL all-gather (sendbuf == recvbuf + rank * count), no staging copy, no allocation."""

    def __init__(self, local_shape, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b = int(local_shape[0])
        self.full = torch.empty((self.world * self.b,) + tuple(local_shape[1:]), dtype=dtype, device=device)
        self.local = self.full[self.rank * self.b:(self.rank + 1) * self.b]
        self._cpu = self.full.device.type == "cpu"

    def all_gather(self, async_op=False):
        import torch.distributed as dist
        # gloo (CPU tests) has no in-place form: it gets a copy of the slice
        return dist.all_gather_into_tensor(self.full, self.local.clone() if self._cpu else self.local, group=self.group, async_op=async_op)


def all_gather_output(local, group=None):
    """[B_local, ...] on every rank -> [world * B_local, ...] on every rank, in place (sendbuf = recvbuf + rank*count).
    Requires torch.distributed to be initialised with the nccl (GPU) or gloo (CPU tests) backend."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    out[rank * local.shape[0]:(rank + 1) * local.shape[0]].copy_(local)
    dist.all_gather_into_tensor(out, out[rank * local.shape[0]:(rank + 1) * local.shape[0]].clone() if local.device.type == "cpu" else
                                out[rank * local.shape[0]:(rank + 1) * local.shape[0]], group=group)
    return out
