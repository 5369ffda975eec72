"""Data-parallel plumbing: one process per GPU (torchrun), gradients of ALL parameters live in one flat
fp32 buffer that is all-reduced with NCCL over NVLink once per optimizer step, in two slices — the
shapelet-expert slice is launched as soon as its last gradient lands (post-accumulate hooks), so it
overlaps the deep expert's backward.  Replaces the reference's nn.DataParallel
(experiment_classification.py:279-281), which cannot gather ModelInfo.

The same class runs on CPU with the gloo backend (tests/test_ddp_gloo.py); nothing here is CUDA specific.
"""
import os

import torch
import torch.distributed as dist


def init_distributed():
    """Reads torchrun's environment.  Returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if backend == "nccl" else None)
    return rank, local, world


class FlatGradAllReduce:
    """Owns a flat gradient buffer; parameter .grad tensors are views into it."""

    def __init__(self, model, world_size, first_slice_prefix="sbm.", overlap=True):
        self.world = world_size
        self.params = [p for p in model.parameters() if p.requires_grad]
        names = {id(p): n for n, p in model.named_parameters()}
        # slice 0: shapelet expert (its backward finishes first: it hangs off the loss through the gate
        # and the auxiliary CE), slice 1: everything else
        first = [p for p in self.params if names[id(p)].startswith(first_slice_prefix)]
        rest = [p for p in self.params if not names[id(p)].startswith(first_slice_prefix)]
        self.slices = [s for s in (first, rest) if s]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=ref.device)
        self.bounds, off = [], 0
        for sl in self.slices:
            beg = off
            for p in sl:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.bounds.append((beg, off))
        self.pending = [0] * len(self.slices)
        self.handles = []
        self.armed = False
        self.overlap = overlap and world_size > 1
        if self.overlap:
            for si, sl in enumerate(self.slices):
                for p in sl:
                    p.register_post_accumulate_grad_hook(self._make_hook(si))
        if world_size > 1:
            self.broadcast_parameters(model)

    def broadcast_parameters(self, model):
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=0)

    def _make_hook(self, si):
        def hook(_param):
            if not self.armed:
                return
            self.pending[si] -= 1
            if self.pending[si] == 0:
                self._launch(si)
        return hook

    def _launch(self, si):
        beg, end = self.bounds[si]
        self.handles.append(dist.all_reduce(self.flat[beg:end], op=dist.ReduceOp.SUM, async_op=True))

    def zero_grad(self):
        self.flat.zero_()

    def arm(self):
        """Call before the backward whose gradients will be reduced (the last micro-step when accumulating)."""
        if self.overlap:
            self.armed = True
            self.pending = [len(sl) for sl in self.slices]

    def finish(self):
        """Wait for / launch the all-reduces and average.  Gradients are then identical on every rank."""
        if self.world <= 1:
            return
        if self.overlap and self.armed:
            for si, left in enumerate(self.pending):
                if left > 0:            # a parameter received no gradient this step
                    self._launch(si)
            self.armed = False
        else:
            self.handles.append(dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True))
        for h in self.handles:
            h.wait()
        self.handles = []
        self.flat.mul_(1.0 / self.world)

    def nbytes(self):
        return self.flat.numel() * 4
