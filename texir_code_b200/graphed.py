"""One optimisation step of the Mat stage as a CUDA graph.

The reference's training step (trainer/train_material.py:436-450) is

    out = self.model(mvp, id, cam_position, stage)      # models/mat_nvdiffrast.py:105-191
    loss = ...(out['rgb'], ground truth)
    loss.backward()
    optimizer.step()

At the training sample count (16 spp, configs/syn.conf:24) the device needs 0.75 ms for the first three lines and the
Python side about as long to issue them -- on a busy host, or with one process per GPU, longer.  GraphedMatStep captures
zero_grad + forward + loss + backward (+ the gradient synchronisation of a sharded model) of stage 0, 1 or 2 ONCE and replays it: one
cudaGraphLaunch per step.  The Cranley-Patterson shift still comes from torch's CPU generator, bit for bit
(sample_util.GraphRng: the engine state is copied in before every replay; the host advances torch's generator itself).

    step = GraphedMatStep(model, stage=2, loss_fn=lambda out, gt: (out['rgb'] - gt).abs().mean(), example=(mvp, cam, gt))
    for it in range(n):
        loss = step(mvp_i, cam_i, gt_i)        # gradients are in model.materials_*.grad
        optimizer.step()                       # (no zero_grad(set_to_none=True) in between: the .grad tensors are the
                                               #  graph's own memory, every replay overwrites them)

What is frozen at capture: tensor shapes, `stage`, sample counts, the model's textures as OBJECTS (their values may change:
the pyramids of the trainable ones are rebuilt inside the graph; the frozen IrT's pyramid and the radiance texture of the
scene are not re-read -- build a new GraphedMatStep after swapping them), and cache_gbuffers (ignored: the G-buffer is part
of the graph).  Construction runs `warmup` eager steps before the capture: the parameters' .grad are overwritten and
torch's CPU generator advances by those steps' draws -- seed after constructing if a run has to be reproducible.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import sample_util


class GraphedMatStep:
    def __init__(self, model, stage: int, loss_fn: Callable, example, sync: Optional[Callable] = None, warmup: int = 3):
        """example = (mvp [V,4,4], cam_position [3], target): tensors of the shapes every later call will have (CUDA);
        loss_fn(result dict of model.forward, target) -> scalar; sync(): gradient synchronisation to run after backward
        inside the graph (e.g. lambda: parallel.allreduce_gradients(model) for a pixel-sharded model)."""
        if stage not in (0, 1, 2):
            raise ValueError("GraphedMatStep covers the optimisation stages 0, 1 and 2")
        dev = model.scene.device
        self.model, self.stage, self.loss_fn, self.sync = model, stage, loss_fn, sync
        self.mvp, self.cam, self.target = (t.detach().to(dev).clone() for t in example)
        self.rng = sample_util.GraphRng(dev)
        self._cached = getattr(model, "cache_gbuffers", False)
        model.cache_gbuffers = False
        try:
            # warm-up on a side stream (allocator, lazy initialisation, NCCL communicators), eager: the generator advances
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            sample_util.finish_pending()
            model.zero_grad(set_to_none=True)     # the capture allocates the gradients inside the graph's pool
            self.graph = torch.cuda.CUDAGraph()
            with self.rng, torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.loss, self.out = self._body()
        finally:
            model.cache_gbuffers = self._cached

    def _body(self):
        # The trainable textures enter the step as FRESH leaves over the same storage.  A parameter's AccumulateGrad node
        # remembers the stream it was created on; one that survives from an earlier eager step (default stream) makes the
        # engine synchronise the capturing stream with the legacy stream, which invalidates the capture.  A fresh leaf gets
        # its node inside the capture, on the capturing stream, whatever ran before.
        model = self.model
        model.zero_grad(set_to_none=True)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        fresh = {n: p.detach().requires_grad_(True) for n, p in named}
        out = torch.func.functional_call(model, fresh, (self.mvp, 0, self.cam, self.stage))
        loss = self.loss_fn(out, self.target)
        loss.backward()
        for n, p in named:
            p.grad = fresh[n].grad
        if self.sync is not None:
            self.sync()
        return loss.detach(), {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}

    def __call__(self, mvp: torch.Tensor, cam_position: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """one step: copies the inputs into the graph's buffers, replays; returns the loss (a tensor the next call
        overwrites); self.out holds the result dict of forward, the parameters' .grad the gradients"""
        self.mvp.copy_(mvp, non_blocking=True)
        self.cam.copy_(cam_position, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        return self.replay()

    def replay(self) -> torch.Tensor:
        """the step on the inputs already in self.mvp / self.cam / self.target"""
        self.rng.before_replay()
        self.graph.replay()
        self.rng.after_replay()
        return self.loss
