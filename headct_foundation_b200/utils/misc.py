"""Hot-path helpers mirrored from `src/utils/misc.py`: MultiCropWrapper (:447-484), the EMA teacher update
(:386-397) and the freeze rule of LoRA / linear-probe fine-tuning (:349-363).  Everything else in the reference's misc.py (meters, checkpoint I/O, plotting) is host glue and out
of scope."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as HF


class MultiCropWrapper(nn.Module):
    """One backbone pass per run of equally-sized crops, head on the CLS token."""

    def __init__(self, backbone, head):
        super().__init__()
        backbone.fc, backbone.head = nn.Identity(), nn.Identity()
        self.backbone = backbone
        self.head = head

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x):
        if not isinstance(x, list):
            x = [x]
        sizes = [inp.shape[-1] for inp in x]
        outs, start = [], 0
        for end in range(1, len(x) + 1):
            if end == len(x) or sizes[end] != sizes[start]:
                _out = self.backbone(torch.cat(x[start:end]))
                if isinstance(_out, tuple):
                    _out = _out[0]
                outs.append(_out)
                start = end
        output = outs[0] if len(outs) == 1 else torch.cat(outs)
        cls_feature = output[:, 0, :]
        return {"dino_output": self.head(cls_feature)}


def set_requires_grad_false(*models, lora: bool = False) -> None:
    """misc.py:349-363.  lora=True keeps the LoRA factors, every bias, the patch / position embeddings and the norms
    trainable (a substring match on the parameter NAME, as the reference does); otherwise everything is frozen.
    Frozen weights cost nothing in our backward: `BlockFn` skips their wgrad GEMMs."""
    for model in models:
        for name, param in model.named_parameters():
            param.requires_grad = bool(lora) and ("lora" in name or "bias" in name or "embeddings" in name or "norm" in name)


@torch.no_grad()
def _update_momentum_encoder(model: nn.Module, momentum_model: nn.Module, m: float) -> None:
    """param_k <- m * param_k + (1 - m) * param_q over zip(parameters()) -- one multi-tensor launch."""
    HF.ema_update(list(model.parameters()), list(momentum_model.parameters()), m)


update_momentum_encoder = _update_momentum_encoder
