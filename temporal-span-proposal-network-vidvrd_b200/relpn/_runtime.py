"""Shared plumbing of the module mirrors: device placement of inputs and weights."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import ops
from ..batch import DeviceBatch, HostBatch


def compute_device(*tensors) -> torch.device:
    """The CUDA device the kernels run on: the inputs' device if they are CUDA tensors, else the
    current device.  The reference's predict.py keeps everything on the CPU (predict.py:21-27):
    CPU inputs are copied to the GPU and the results copied back, so the call is drop-in."""
    ops.require_device()
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


class WeightCache:
    """fp32 contiguous device copies of a module's parameters, refreshed when they change."""

    def __init__(self):
        self._key = None
        self._dev: Dict[str, torch.Tensor] = {}

    def get(self, module: torch.nn.Module, device: torch.device) -> Dict[str, torch.Tensor]:
        sd = {k: v for k, v in module.state_dict().items()}
        key = (str(device),) + tuple((k, v.data_ptr(), v._version, tuple(v.shape)) for k, v in sd.items())
        if key != self._key:
            self._dev = {k: v.detach().to(device, torch.float32).contiguous() for k, v in sd.items()}
            self._key = key
        return self._dev


def batch_from_pair_lists(pair_list, device, need_motion: bool) -> DeviceBatch:
    """Pack the tracklet fields of a list of PairList into one device batch."""
    boxes, span, cls, motion = [], [], [], []
    for pl in pair_list:
        cl = pl.get_field("track_cls_logits")
        n = int(cl.shape[0])
        if pl.has_field("boxes"):
            boxes.append(pl.get_field("boxes"))
            span.append(pl.get_field("span"))
        else:       # reference-mode PairList: no geometry, a 1-frame placeholder keeps the layout valid
            boxes.append(torch.zeros((n, 1, 4)))
            span.append(torch.tensor([[0, 1]] * n, dtype=torch.int32).reshape(n, 2))
        cls.append(cl)
        if need_motion:
            motion.append(pl.get_field("motion"))
    host = HostBatch(boxes, span, cls, motion if need_motion else None)
    return host.to_device(device)


def like_input(t: Optional[torch.Tensor], ref_is_cuda: bool) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t if ref_is_cuda else t.cpu()
