"""Batch-sharded data parallelism: one process per GPU, ONE NCCL all-reduce (sum) over the gradient arena (all models' flat
gradient buffers back to back, nn.unify_gradients) between backward and the fused Adam launch; the 1/world_size average is folded into the Adam kernel's gradient scale.

The reference has no multi-GPU code at all (SURVEY.md 2.1 "Parallelism strategies present in the reference: none");
every operator on the path is per-image (no BatchNorm), so sharding the raw batch by rank is exact: mean-type losses
(CE, MSE) average over ranks. torch.distributed is used for rendezvous and the NCCL call only.
"""
import torch
import torch.distributed as dist


class GradSync:
    """Callable hook for ManipulationClassification.training_step_device(grad_sync=...)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.gscale = 1.0 / self.world

    def __call__(self, stores):
        if self.world == 1:
            return
        arena = getattr(stores[0], 'arena', None)
        # stores laid out in one arena (nn.unify_gradients): ONE collective over the whole bucket
        if arena is not None and all(getattr(s, 'arena', None) is arena for s in stores) \
                and sum(s.gflat.numel() for s in stores) == arena.numel():
            dist.all_reduce(arena, op=dist.ReduceOp.SUM, group=self.group)
            return
        for s in stores:
            dist.all_reduce(s.gflat, op=dist.ReduceOp.SUM, group=self.group)


def shard_batch(batch, rank, world):
    """Rank r takes raw patches [r*B/W, (r+1)*B/W); the class-major manipulation stack stays
    self-consistent per rank because labels are generated from the LOCAL batch size."""
    b = batch.shape[0]
    if b % world:
        raise ValueError('global batch {} is not divisible by world size {}'.format(b, world))
    per = b // world
    return batch[rank * per:(rank + 1) * per]


def broadcast_parameters(stores, src=0, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for s in stores:
            dist.broadcast(s.flat, src=src, group=group)
