"""Training-step drivers for the drop-in models.

`eager_step` is the per-batch body of the reference trainer (FoodRec/common/trainer.py:177-224):
zero_grad, `calculate_loss`, sum, backward, optimizer step.  `GraphedTrainStep` runs the same body
as ONE CUDA-graph replay: the step is ~100 small launches (the propagation kernels are tens of
microseconds each at Allrecipes scale), so without a graph it is bound by CPU launch latency, not
by the GPU.  Numerics are identical to the eager body (same kernels, same order).
"""
from __future__ import annotations

import torch


def mark_parameters_updated(model):
    """Invalidate everything cached from the parameter values (propagated evaluation tables, SCHGN's
    user-independent scorer tables).  Needed because a CUDA-graph replay rewrites the parameters in
    place without going through ATen: neither `data_ptr` nor `_version` changes."""
    model._param_generation = getattr(model, "_param_generation", 0) + 1
    if getattr(model, "_eval_cache", None) is not None:
        model._eval_cache = None


def eager_step(model, optimizer, batch, grad_hook=None):
    optimizer.zero_grad()
    losses = model.calculate_loss(batch)
    loss = sum(losses) if isinstance(losses, tuple) else losses
    loss.backward()
    if grad_hook is not None:
        grad_hook(model)
    optimizer.step()
    mark_parameters_updated(model)
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


class FusedAdam(torch.optim.Optimizer):
    """`torch.optim.Adam(params, lr, betas, eps)` (amsgrad = False, weight_decay = 0: the reference's optimizer,
    FoodRec/common/trainer.py:144) as ONE multi-tensor launch of this library (`fr_adam_step`): dense update of every
    parameter that has a gradient, step counter on the device, capturable in a CUDA graph.  State keys (`exp_avg`,
    `exp_avg_sq`) are those of torch's Adam, so optimizer checkpoints carry over except for `step`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._dev_state = {}

    @torch.no_grad()
    def step(self, closure=None):
        from . import _lib
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _lib.FoodRecError("FusedAdam needs CUDA parameters (foodrec_b200 has no CPU path)")
            ds = self._dev_state.get(gi)
            if ds is None:
                ds = self._dev_state[gi] = (torch.zeros(1, dtype=torch.int32, device=dev),
                                            torch.zeros(2, dtype=torch.float32, device=dev))
            tensors = (_lib.AdamTensor * len(ps))()
            for t, p in zip(tensors, ps):
                if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise _lib.FoodRecError("FusedAdam handles contiguous fp32 parameters with dense gradients")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                t.param, t.grad, t.exp_avg, t.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                t.n = p.numel()
            b1, b2 = group["betas"]
            _lib.check(_lib.lib.fr_adam_step(tensors, len(ps), float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                             ds[0].data_ptr(), ds[1].data_ptr(), _lib.stream_ptr()), "fr_adam_step")
        return loss


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
            # every term in one device vector (inside the graph), so a trainer reads them with ONE device->host copy
            self.loss_vec = torch.stack([l.detach().reshape(()).float() for l in self.losses])

    def __call__(self, batch: dict):
        """`batch` tensors may live on the device or in (pinned) host memory: they are copied straight into the
        static inputs of the graph."""
        for k in self.keys:
            self.static[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        mark_parameters_updated(self.model)
        return self.losses

    def loss_values(self):
        """The loss terms of the last replay as python floats: one synchronising device->host copy for all of them
        (the reference trainer's per-term `.item()`, trainer.py:186, costs one round trip per term)."""
        return self.loss_vec.tolist()


class DeviceBatchSampler:
    """Training batches assembled on the device: a shuffled pass over the training interactions with one
    uniformly drawn negative per sample (`fr_sample_negatives`), as `TrainDataLoader` + `DataLoader(shuffle=True)`
    produce them (FoodRec/utils/dataloader.py:50-77,145-151) -- without the per-sample python rejection loop and
    without a host round trip.  Negatives avoid the user's training items and, when the dataset carries them,
    their validation / test items (`validTestRatings`; else `validRatings` mapped through `valid_users`, and
    `testRatings`), like the reference.  Same distribution, a
    different random stream: parity is statistical."""

    def __init__(self, dataset, batch_size: int, device, seed: int = 0, drop_last: bool = False):
        import numpy as np
        coo = dataset.train_coo_matrix
        self.n_items, self.batch_size, self.device = int(dataset.n_items), int(batch_size), torch.device(device)
        u = np.asarray(coo.row, dtype=np.int64)
        i = np.asarray(coo.col, dtype=np.int64)
        self.users = torch.from_numpy(u).to(self.device)
        self.items = torch.from_numpy(i).to(self.device)
        eu, ei = [u], [i]
        n_users = int(dataset.n_users)
        vt = getattr(dataset, "validTestRatings", None)
        if vt is not None:
            # the reference's own exclusion structure: dict user -> set of held-out items
            # (FoodRec/utils/dataset.py:35,93-113, read by utils/dataloader.py:145-151)
            us = np.fromiter((uu for uu, items in vt.items() for _ in items), dtype=np.int64)
            it = np.fromiter((j for items in vt.values() for j in items), dtype=np.int64)
            eu.append(us)
            ei.append(it)
        else:
            # `validRatings` is POSITIONAL in the reference (paired with `valid_users`: users without validation
            # rows are skipped, dataset.py:32,115-135); `testRatings` has one list per user
            for name, owners in (("validRatings", "valid_users"), ("testRatings", "test_users")):
                lists = getattr(dataset, name, None)
                if lists is None:
                    continue
                ids = getattr(dataset, owners, None)
                if ids is None:
                    if len(lists) != n_users:
                        raise ValueError(f"dataset.{name} has {len(lists)} lists for {n_users} users and no "
                                         f"dataset.{owners} to map them: cannot attribute held-out items")
                    ids = np.arange(n_users, dtype=np.int64)
                ids = np.asarray(ids, dtype=np.int64)
                if ids.shape[0] != len(lists):
                    raise ValueError(f"dataset.{owners} and dataset.{name} differ in length")
                lens = np.fromiter((len(x) for x in lists), dtype=np.int64, count=len(lists))
                eu.append(np.repeat(ids, lens))
                ei.append(np.fromiter((j for x in lists for j in x), dtype=np.int64, count=int(lens.sum())))
        keys = np.unique(np.concatenate(eu) * self.n_items + np.concatenate(ei))
        ptr = np.zeros(int(dataset.n_users) + 1, dtype=np.int64)
        np.cumsum(np.bincount(keys // self.n_items, minlength=int(dataset.n_users)), out=ptr[1:])
        self.excl_ptr = torch.from_numpy(ptr).to(self.device)
        self.excl_idx = torch.from_numpy((keys % self.n_items).astype(np.int32)).to(self.device)
        self.seed, self.drop_last, self._step = int(seed), drop_last, 0
        self._fail = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(self.seed)

    def __len__(self):
        n = self.users.numel()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def negatives(self, users: torch.Tensor) -> torch.Tensor:
        from . import _lib
        users = users.to(self.device, torch.int64).contiguous()
        out = torch.empty_like(users)
        self._step += 1
        _lib.check(_lib.lib.fr_sample_negatives(
            self.excl_ptr.data_ptr(), self.excl_idx.data_ptr(), users.data_ptr(), users.numel(), self.n_items,
            self.seed & (2 ** 64 - 1), self._step, out.data_ptr(), self._fail.data_ptr(), _lib.stream_ptr()),
            "fr_sample_negatives")
        return out

    def failures(self) -> int:
        """Samples that found no admissible item so far (a user who interacted with almost every item)."""
        return int(self._fail.item())

    def __iter__(self):
        perm = torch.randperm(self.users.numel(), device=self.device, generator=self._gen)
        for b in range(len(self)):
            sel = perm[b * self.batch_size:(b + 1) * self.batch_size]
            u = self.users[sel]
            yield {"u_id": u, "pos_i_id": self.items[sel], "neg_i_id": self.negatives(u)}
