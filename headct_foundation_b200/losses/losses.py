"""DINOLoss on B200 kernels.  Drop-in for `src/losses/losses.py:46-102`."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import functional as HF


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs, nepochs,
                 student_temp=0.1, center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp))

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, student_output, teacher_output, epoch):
        temp = float(self.teacher_temp_schedule[epoch])
        with torch.autocast(device_type="cuda", enabled=False):
            loss = HF.DinoLossFn.apply(student_output, teacher_output.detach(), self.center, self.ncrops,
                                       self.student_temp, temp)
            self.update_center(teacher_output)
        return loss

    @torch.no_grad()
    def update_center(self, teacher_output):
        # in place on the registered buffer (the reference rebinds self.center; values are identical)
        HF.center_update(self.center, teacher_output.detach(), self.center_momentum)
