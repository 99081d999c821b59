"""CUDA-graph replay of launch-bound forwards and training steps.

Feature extraction at small batch (`notebooks/extract_feature_sample.ipynb` cell 12: `model(x)` under no_grad; BASELINE
config 5 sweeps batch 1-512) enqueues ~90 kernels of a few microseconds each through the C ABI, so below batch ~16 the
host, not the GPU, sets the pace.  `GraphedForward` captures one no-grad forward of a drop-in module for a fixed input
shape into a CUDA graph (the kernels are enqueued on torch's capture stream through the same C-ABI calls -- the
library never synchronises or allocates) and replays it with one launch per call.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import torch


def _tree_map(fn, obj):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_tree_map(fn, o) for o in obj)
    if isinstance(obj, dict):
        return {k: _tree_map(fn, v) for k, v in obj.items()}
    return obj


class GraphedForward:
    """`y = GraphedForward(model, example)(x)`: same values as `model(x)` under `torch.no_grad()`.

    * `x` must have the example's shape / dtype / device; it is copied into a static input buffer.
    * The returned tensors are the graph's static output buffers: they are overwritten by the next call
      (`clone=True` hands back copies instead).
    * The graph holds the parameters' bf16 GEMM copies by address; when any parameter's version counter has moved
      (optimizer step, `load_state_dict`) the next call re-captures.
    """

    def __init__(self, module: torch.nn.Module, example: torch.Tensor, warmup: int = 2, clone: bool = False):
        if not example.is_cuda:
            raise RuntimeError("GraphedForward needs a CUDA example input (no CPU fallback)")
        self.module, self.clone, self.warmup = module, clone, warmup
        self.static_in = example.detach().clone()
        self.graph = None
        self.static_out: Any = None
        self._versions: Tuple[int, ...] = ()
        self._capture()

    def _param_versions(self) -> Tuple[int, ...]:
        return tuple(p._version for p in self.module.parameters())

    def _capture(self) -> None:
        dev = self.static_in.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():         # warm-up off the capture: bf16 weight copies, func attributes
            for _ in range(max(1, self.warmup)):
                self.module(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.module(self.static_in)
        self._versions = self._param_versions()

    def __call__(self, x: torch.Tensor):
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"GraphedForward was captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        if self._param_versions() != self._versions:
            self._capture()
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return _tree_map(torch.clone, self.static_out) if self.clone else self.static_out


class GraphedTrainStep:
    """One whole training step -- zero_grad, forward, loss, backward, per-parameter clip + AdamW -- as ONE graph launch.

    Enqueueing the ~490 kernels of an MAE step through Python / ctypes costs the host ~13 ms, which caps small per-GPU
    batches (13 ms of GPU work at batch 32); the reference pays the same kind of cost through eager PyTorch plus 254
    `.item()` syncs (engine_pretrain_mae.py:52-86, misc.py:374-383).  Usage:

        step = GraphedTrainStep(model, FusedAdamW(...), example_batch)     # `model(x)` must return the loss first
        for x in loader:
            optimizer.param_groups[0]["lr"] = schedule(it)                 # read at replay time (device-side scalars)
            loss = step(x)                                                 # static loss tensor, read it before the next call

    Everything the step needs per iteration lives in device memory (`FusedAdamW.step_captured` / `advance`); the mask
    noise comes from torch's CUDA generator, which is graph-aware.

    Data parallel (`sync_grads`, default: on when torch.distributed runs with more than one rank): the gradient all-reduce
    (mean over ranks, `parallel.allreduce_mean_grads_`: coalesced NCCL groups, captured like any other kernel) sits INSIDE
    the graph -- per ~64 MB bucket on a forked communication stream as soon as backward has produced the bucket's gradients
    (`overlap_grad_sync`, default), joined before the update --, so a small per-GPU batch (global batch 256 over 8 GPUs = 32 per GPU, 13 ms of
    GPU work) is not capped by the host's ~13 ms of launch work per step.  Pass the bare module, not a DistributedDataParallel
    wrapper (its reducer hooks cannot be captured); parameters are broadcast from rank 0 first, as DDP would.  The graph
    holds NCCL kernels: drop the step (`del step`) before `torch.distributed.destroy_process_group()`, or that call waits
    forever.  Checked on two ranks against the eager backward -> all-reduce -> update sequence
    (tests/test_gpu_parity_full.py::test_graphed_step_with_captured_allreduce: ranks bit-identical after three steps).
    """

    def __init__(self, model: torch.nn.Module, optimizer, example: torch.Tensor, loss_of=None, warmup: int = 3,
                 sync_grads: Optional[bool] = None, overlap_grad_sync: bool = True, bucket_bytes: int = 64 << 20):
        if not example.is_cuda:
            raise RuntimeError("GraphedTrainStep needs a CUDA example input (no CPU fallback)")
        if not hasattr(optimizer, "step_captured"):
            raise TypeError("GraphedTrainStep needs headct_foundation_b200.optim.FusedAdamW")
        from .. import parallel
        if isinstance(model, torch.nn.parallel.DistributedDataParallel):
            raise TypeError("GraphedTrainStep takes the bare module (it reduces the gradients itself inside the graph)")
        self.sync_grads = (parallel.is_dist() and parallel.world()[1] > 1) if sync_grads is None else bool(sync_grads)
        if self.sync_grads:
            parallel.broadcast_params_(model)
        self.model, self.optimizer = model, optimizer
        self.loss_of = loss_of or (lambda out: out[0] if isinstance(out, (tuple, list)) else out)
        self.static_in = example.detach().clone()
        dev = example.device
        # eager warm-up (optimizer state, bf16 weight copies, allocator, lazy kernel attributes) on a side stream; the
        # parameters and the optimizer state it touches are put back afterwards, so capturing does not train the model
        params = [p for g in optimizer.param_groups for p in g["params"]]
        saved_p = [p.detach().clone() for p in params]
        saved_s = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in optimizer.state[p].items()} if p in optimizer.state
                   and len(optimizer.state[p]) else None for p in params]
        # Gradient exchange under backward (DDP's overlap, in capturable form): the parameters are bucketed in reverse
        # registration order (~the order backward produces their gradients); a post-accumulate hook on every parameter counts
        # its bucket down, and the hook that completes a bucket forks the communication stream off the stream backward
        # runs on and issues that bucket's coalesced all-reduce there.  Captured, the fork / join become graph edges: the
        # all-reduce of a bucket runs beside the backward kernels of the layers below it.
        self._overlap = bool(self.sync_grads and overlap_grad_sync)
        hooks = []
        if self._overlap:
            comm = torch.cuda.Stream(device=dev)
            buckets, size = [[]], 0
            for p in reversed([q for q in params if q.requires_grad]):
                nbytes = p.numel() * 4
                if buckets[-1] and size + nbytes > bucket_bytes:
                    buckets.append([])
                    size = 0
                buckets[-1].append(p)
                size += nbytes
            pending = [0] * len(buckets)

            def arm():
                for i, b in enumerate(buckets):
                    pending[i] = len(b)

            def make_hook(bi):
                def hook(_p):
                    pending[bi] -= 1
                    if pending[bi] == 0:
                        comm.wait_stream(torch.cuda.current_stream(dev))
                        with torch.cuda.stream(comm):
                            parallel.allreduce_mean_grads_(buckets[bi], bucket_bytes=1 << 62)
                return hook
            for bi, b in enumerate(buckets):
                for p in b:
                    hooks.append(p.register_post_accumulate_grad_hook(make_hook(bi)))

            def sync_after_backward():
                cur = torch.cuda.current_stream(dev)
                late = [p for i, b in enumerate(buckets) if pending[i] > 0 for p in b]     # a bucket with a gradient-less tensor
                cur.wait_stream(comm)
                if late:
                    parallel.allreduce_mean_grads_(late)
        else:
            def arm():
                pass

            def sync_after_backward():
                if self.sync_grads:
                    parallel.allreduce_mean_grads_(params)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                optimizer.zero_grad(set_to_none=True)
                arm()
                self.loss_of(model(self.static_in)).backward()
                sync_after_backward()                  # also brings the NCCL communicator up before the capture
                optimizer.step()
            with torch.no_grad():
                for p, sp, ss in zip(params, saved_p, saved_s):
                    p.copy_(sp)
                    st = optimizer.state.get(p)
                    if st:
                        for k, v in st.items():
                            if torch.is_tensor(v):
                                v.copy_(ss[k]) if ss is not None else v.zero_()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        del saved_p, saved_s
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        optimizer.prepare_capture()
        # With collectives inside, other threads of this process touch the CUDA API while the capture runs (the NCCL
        # watchdog polls its events): "thread_local" keeps the capture's legality checks to the capturing thread.
        if self.sync_grads:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local" if self.sync_grads else "global"):
            loss = self.loss_of(model(self.static_in))
            arm()
            loss.backward()
            sync_after_backward()
            optimizer.step_captured()
        for h in hooks:                                    # the graph keeps the exchange; eager use of the model is unaffected
            h.remove()
        self.static_loss = loss.detach()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"GraphedTrainStep was captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        self.static_in.copy_(x, non_blocking=True)
        self.optimizer.advance()
        self.graph.replay()
        return self.static_loss
