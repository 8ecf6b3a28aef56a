"""Training-step drivers for the drop-in models.

`eager_step` is the per-batch body of the reference trainer (FoodRec/common/trainer.py:177-224):
zero_grad, `calculate_loss`, sum, backward, optimizer step.  `GraphedTrainStep` runs the same body
as ONE CUDA-graph replay: the step is ~100 small launches (the propagation kernels are tens of
microseconds each at Allrecipes scale), so without a graph it is bound by CPU launch latency, not
by the GPU.  Numerics are identical to the eager body (same kernels, same order).
"""
from __future__ import annotations

import torch


def eager_step(model, optimizer, batch, grad_hook=None):
    optimizer.zero_grad()
    losses = model.calculate_loss(batch)
    loss = sum(losses) if isinstance(losses, tuple) else losses
    loss.backward()
    if grad_hook is not None:
        grad_hook(model)
    optimizer.step()
    return losses


def allreduce_mean_grads(group=None):
    """Gradient hook for data-parallel replicas: one NCCL all-reduce per dense gradient (capturable)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)

    def hook(model):
        for p in model.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad, group=group)
                p.grad.div_(world)
    return hook


class OverlappedGradAllReduce:
    """Data-parallel gradient averaging that overlaps the backward: every parameter's gradient is
    all-reduced (NCCL, AVG) on a communication stream as soon as autograd has accumulated it, so the
    user-table gradient (ready after the user-item backward) travels while the item-side graphs are
    still back-propagating.  Call the object between `backward()` and `optimizer.step()` to join the
    streams.  Works under CUDA-graph capture (the communication stream becomes a parallel branch)."""

    def __init__(self, model, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        dev = next(model.parameters()).device
        self.comm = torch.cuda.Stream(device=dev)
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in model.parameters() if p.requires_grad]

    def _hook(self, p):
        cur = torch.cuda.current_stream()
        self.comm.wait_stream(cur)
        with torch.cuda.stream(self.comm):
            self.dist.all_reduce(p.grad, op=self.dist.ReduceOp.AVG, group=self.group)

    def __call__(self, model=None):
        torch.cuda.current_stream().wait_stream(self.comm)

    def remove(self):
        for h in self.handles:
            h.remove()


class GraphedTrainStep:
    """Capture `zero_grad -> calculate_loss -> backward -> optimizer.step` once, replay per batch.

    `example_batch` fixes the batch shapes (the reference's DataLoader yields a short last batch:
    run that one through `eager_step`).  The optimizer must be capturable (e.g.
    `torch.optim.Adam(..., capturable=True)`).  `__call__(batch)` copies the batch tensors (host
    pinned or device) into the static inputs, replays, and returns the tuple of loss tensors
    (device, overwritten by the next call).
    """

    def __init__(self, model, optimizer, example_batch: dict, keys=None, warmup: int = 3, grad_hook=None):
        self.model, self.optimizer = model, optimizer
        dev = next(model.parameters()).device
        self.keys = list(keys) if keys is not None else list(example_batch.keys())
        self.static = {k: torch.empty_like(example_batch[k], device=dev).copy_(example_batch[k]) for k in self.keys}
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):  # allocates lazily-created workspaces and optimizer state outside the graph
                eager_step(model, optimizer, self.static, grad_hook)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            losses = model.calculate_loss(self.static)
            loss = sum(losses) if isinstance(losses, tuple) else losses
            loss.backward()
            if grad_hook is not None:
                grad_hook(model)
            optimizer.step()
        self.losses = losses if isinstance(losses, tuple) else (losses,)
        self.loss_vec = None

    def __call__(self, batch: dict):
        for k in self.keys:
            self.static[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.losses
