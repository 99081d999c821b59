"""Train-step glue as multi-tensor kernels.

`FusedAdamW` reproduces what the reference engines do between backward and the next forward --
`clip_gradients` (per-parameter L2 clip, src/utils/misc.py:374-383) followed by `torch.optim.AdamW.step`
(src/utils/optimizers.py:354-360) -- in two launches over all parameters, with no host sync
(the reference pays one `.item()` per parameter tensor).  State layout (`exp_avg`, `exp_avg_sq`, `step`)
matches torch.optim.AdamW so optimizer checkpoints interchange.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._cabi import call, stream_ptr
from .functional import mark_updated, shadow_of


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, clip_grad: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_grad=clip_grad)
        super().__init__(params, defaults)
        self._tables = {}
        self._graph_state = {}

    def _steps(self, ps):
        """Per-parameter step counts BEFORE this update (state['step'] is a CPU tensor, as in torch.optim.AdamW)."""
        return [int(self.state[p]["step"]) for p in ps]

    def _table(self, gi, group):
        ps = [p for p in group["params"] if p.grad is not None]
        rows = []
        shadows = []
        for p in ps:
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("FusedAdamW expects fp32 CUDA parameters and gradients")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if g is not p.grad:
                p.grad = g
            sh = shadow_of(p)        # bf16 GEMM-operand copy, rewritten by the AdamW launch itself
            shadows.append(sh)
            rows.append([p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                         0 if sh is None else sh.data_ptr(), 0])
        # column 6: how many steps each tensor is behind the first one (torch bias-corrects per parameter; a tensor whose
        # grad was None for a while -- cancel_gradients_last_layer -- lags).  Part of the signature: the table is rebuilt
        # whenever the set of participating tensors or their relative step counts change.
        steps = self._steps(ps)
        for row, s in zip(rows, steps):
            row[6] = steps[0] - s
        sig = tuple(tuple(r) for r in rows)
        hit = self._tables.get(gi)
        if hit is None or hit[0] != sig:
            dev = ps[0].device
            table = torch.tensor(rows, dtype=torch.int64).to(dev)
            norms = torch.empty(len(rows), dtype=torch.float32, device=dev)
            hit = (sig, table, norms)
            self._tables[gi] = hit
        return ps, hit[1], hit[2], shadows, steps[0] + 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            ps, table, norms, shadows, step = self._table(gi, group)
            for p in ps:
                self.state[p]["step"] += 1
            st = stream_ptr(ps[0].device)
            clip = float(group.get("clip_grad", 0.0) or 0.0)
            if clip > 0:
                call("hct_grad_norms_multi", table.data_ptr(), table.shape[0], norms.data_ptr(), st)
            b1, b2 = group["betas"]
            call("hct_adamw_multi", table.data_ptr(), table.shape[0], norms.data_ptr(), clip, float(group["lr"]),
                 float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), step, st)
            mark_updated(ps, shadows)     # the kernel wrote parameters (and their bf16 copies) through raw pointers
        return loss

    # ------------------------------------------------------------------ CUDA-graph support (utils/graphs.py)
    def _hyper_values(self, group, step: int):
        b1, b2 = group["betas"]
        return [float(group["lr"]), float(group["weight_decay"]), 1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5, float(step)]

    def prepare_capture(self):
        """Allocate the buffers the captured update reads (pointer table, norm workspace, per-step scalars) OUTSIDE the
        capture: memory allocated while capturing belongs to the graph's pool and its initialising fill would be replayed,
        wiping the scalars staged by `advance()`."""
        for gi, group in enumerate(self.param_groups):
            n = len(group["params"])
            dev = group["params"][0].device
            self._graph_state[gi] = dict(table_host=torch.zeros((n, 7), dtype=torch.int64).pin_memory(),
                                         table=torch.zeros((n, 7), dtype=torch.int64, device=dev),
                                         norms=torch.zeros(n, dtype=torch.float32, device=dev),
                                         hyper=torch.zeros(5, dtype=torch.float32, device=dev), params=[], shadows=[])

    @torch.no_grad()
    def step_captured(self):
        """Enqueue the update with every per-step scalar read from device memory (`hct_adamw_multi_dev`): call it inside a
        CUDA-graph capture (after `prepare_capture()`), then `advance()` before each replay.  The pointer table reaches
        the device through a captured copy from pinned memory (the gradients' addresses are only known while capturing)."""
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            gs = self._graph_state.get(gi)
            if gs is None:
                raise RuntimeError("call FusedAdamW.prepare_capture() before capturing")
            dev = ps[0].device
            rows, shadows = [], []
            for p in ps:
                st = self.state[p]
                if len(st) == 0:
                    raise RuntimeError("run at least one eager step before capturing (optimizer state must exist)")
                if not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW.step_captured needs contiguous gradients")
                sh = shadow_of(p)
                shadows.append(sh)
                rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel(),
                             0 if sh is None else sh.data_ptr(), 0])
            steps = self._steps(ps)          # the lags are constant across replays: every captured tensor steps together
            for row, s in zip(rows, steps):
                row[6] = steps[0] - s
            n = len(rows)
            gs["table_host"][:n].copy_(torch.tensor(rows, dtype=torch.int64))
            gs["params"], gs["shadows"] = ps, shadows
            gs["table"].copy_(gs["table_host"], non_blocking=True)        # captured H2D node from pinned memory
            st = stream_ptr(dev)
            clip = float(group.get("clip_grad", 0.0) or 0.0)
            if clip > 0:
                call("hct_grad_norms_multi", gs["table"].data_ptr(), n, gs["norms"].data_ptr(), st)
            b1, b2 = group["betas"]
            call("hct_adamw_multi_dev", gs["table"].data_ptr(), n, gs["norms"].data_ptr(), clip,
                 gs["hyper"].data_ptr(), float(b1), float(b2), float(group["eps"]), st)

    @torch.no_grad()
    def advance(self):
        """Host side of one replayed step: bump the step counters, stage lr / weight decay / bias corrections for the
        captured launch (async copy on the current stream, ahead of the replay) and re-key the bf16 weight copies."""
        for gi, group in enumerate(self.param_groups):
            gs = self._graph_state.get(gi)
            if gs is None or not gs["params"]:
                continue
            ps = gs["params"]
            step = int(self.state[ps[0]]["step"].item()) + 1
            for p in ps:
                self.state[p]["step"] += 1
            # a fresh pinned staging tensor per step: torch's host allocator keeps it alive until the copy has run, so the
            # host may run ahead of the GPU by several steps without overwriting values still in flight
            staged = torch.tensor(self._hyper_values(group, step), dtype=torch.float32).pin_memory()
            gs["hyper"].copy_(staged, non_blocking=True)
            mark_updated(ps, gs["shadows"])
