"""On-GPU HU windowing.  Replaces the CPU/MONAI `MultipleWindowScaleStack` (`src/data/transforms.py:8-36`) for
volumes that already live on the device; the rest of the MONAI chain (I/O, resampling, crops) is out of scope."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch

from .. import functional as HF

DEFAULT_WINDOWS: Tuple[Tuple[int, int], ...] = ((40, 80), (80, 200), (600, 2800))     # transforms.py:130


class MultipleWindowScaleStack:
    def __init__(self, keys: List[str], window_sizes: List[Tuple[int, int]] = list(DEFAULT_WINDOWS),
                 out_dtype: torch.dtype = torch.float32) -> None:
        self.keys = keys
        self.window_sizes = [tuple(w) for w in window_sizes]
        self.out_dtype = out_dtype

    def __call__(self, data: Dict[str, Any]) -> Dict[str, Any]:
        d = dict(data)
        d["image"] = HF.window_scale_stack(d["image"], self.window_sizes, self.out_dtype)
        return d
