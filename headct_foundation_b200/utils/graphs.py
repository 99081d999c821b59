"""CUDA-graph replay of a launch-bound forward.

Feature extraction at small batch (`notebooks/extract_feature_sample.ipynb` cell 12: `model(x)` under no_grad; BASELINE
config 5 sweeps batch 1-512) enqueues ~90 kernels of a few microseconds each through the C ABI, so below batch ~16 the
host, not the GPU, sets the pace.  `GraphedForward` captures one no-grad forward of a drop-in module for a fixed input
shape into a CUDA graph (the kernels are enqueued on torch's capture stream through the same C-ABI calls -- the
library never synchronises or allocates) and replays it with one launch per call.
"""
from __future__ import annotations

from typing import Any, Tuple

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
