"""Host-facing steady-state loop of the pair stage: pinned host batches in, pinned host results out.

The reference's driver (lib/modeling/predict.py:42-57) feeds one ``PairList`` at a time from a
``DataLoader`` and reads the outputs back with ``.numpy()``.  On a B200 a per-video call is
launch- and PCIe-latency-bound, so the serving loop works on *batches of videos* and keeps
``depth`` of them in flight over three streams:

    h2d stream    :  H2D(i+1) ......................
    compute stream:  kernels(i)  (one CUDA-graph launch, see pipeline.GraphedStage); two compute streams
                     alternate, so the end of step i overlaps the start of step i+1
    d2h stream    :  D2H(i-1) ......................

Every slot owns its device input buffers, its captured graphs (and therefore its output buffers) and
its pinned host result buffers, so nothing is allocated in steady state.  With ``group`` the per-video
triplet records are all-gathered across ranks after the kernels of each step (the one collective of
the path, sharding.py) on a stream of their own: only the D2H copy of the gathered records waits for it.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .batch import DeviceBatch, HostBatch
from .pipeline import GraphedStage, PairStage


class _Slot:
    def __init__(self, stage: PairStage, template: HostBatch, device, graphs: bool, single_graph: bool = True):
        self.batch: DeviceBatch = template.to_device(device, non_blocking=False)
        torch.cuda.synchronize(device)
        self.graphed: Optional[GraphedStage] = stage.capture(self.batch, single=single_graph) if graphs else None
        self.h2d_done = torch.cuda.Event()
        self.kernels_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.host_out: Optional[Dict[str, torch.Tensor]] = None
        self.keep = None
        self.busy = False


class PipelinedStage:
    """``submit(host_batch)`` enqueues H2D -> kernels -> D2H for one batch and returns a ticket;
    ``wait(ticket)`` blocks until that batch's results are in pinned host memory and returns them
    (valid until ``depth`` further submits).  All batches must have the per-video shapes of
    ``template`` (the graphs are captured for them)."""

    def __init__(self, stage: PairStage, template: HostBatch, device="cuda", depth: int = 2, graphs: bool = True,
                 group=None, compute_streams: int = 2, single_graph: bool = True):
        self.stage, self.device, self.depth, self.group = stage, torch.device(device), int(depth), group
        self.main = torch.cuda.current_stream(self.device)
        # Two compute streams, used alternately: the latency-bound tail of step i (feature rows, heads,
        # records - it leaves most of the HBM bandwidth idle) runs underneath the HBM-bound geometry kernel
        # of step i+1.  Slots never share buffers, so the only ordering needed is per slot (events below).
        self.compute = [torch.cuda.Stream(self.device) for _ in range(2 if int(depth) > 1 and compute_streams > 1 else 1)]
        self.s_h2d, self.s_d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        # The collective has its own stream: only the D2H of the gathered records waits for it.  On a compute stream
        # it would hold up the next step behind an NCCL kernel that cannot co-reside with the persistent pair
        # kernel of the step running on the other compute stream.
        self.s_comm = torch.cuda.Stream(self.device) if group is not None else None
        self.slots: List[_Slot] = [_Slot(stage, template, self.device, graphs, single_graph) for _ in range(self.depth)]
        self._next = 0
        self._gathered: List[Optional[torch.Tensor]] = [None] * self.depth

    # ------------------------------------------------------------------------------------------
    def submit(self, host: HostBatch) -> int:
        i = self._next
        self._next = (i + 1) % self.depth
        slot = self.slots[i]
        if slot.busy:
            raise RuntimeError("PipelinedStage: %d batches already in flight; wait() for a ticket first" % self.depth)
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(slot.kernels_done)     # the slot's previous kernels have consumed its inputs
            slot.batch.copy_from(host)
            slot.h2d_done.record(self.s_h2d)
        main = self.compute[i % len(self.compute)]
        with torch.cuda.stream(main):
            main.wait_event(slot.h2d_done)
            main.wait_event(slot.d2h_done)               # the slot's previous results have left the device
            res = slot.graphed.replay() if slot.graphed is not None else self.stage.forward(slot.batch)
            outs = res.host_outputs()
            slot.kernels_done.record(main)
        work = None
        if self.group is not None:                       # the one collective: top-K triplet records
            import torch.distributed as dist
            world = dist.get_world_size(self.group)
            if self._gathered[i] is None:
                self._gathered[i] = torch.empty((world,) + tuple(res.records.shape), dtype=res.records.dtype,
                                                device=self.device)
            with torch.cuda.stream(self.s_comm):
                self.s_comm.wait_event(slot.kernels_done)
                work = dist.all_gather_into_tensor(self._gathered[i], res.records, group=self.group, async_op=True)
            outs["records_all_ranks"] = self._gathered[i]
        if slot.host_out is None:
            slot.host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in outs.items()}
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(slot.kernels_done)
            if work is not None:
                work.wait()                              # this stream waits for the all-gather, nobody else does
            for k, src in outs.items():
                if slot.graphed is None:
                    src.record_stream(self.s_d2h)
                slot.host_out[k].copy_(src, non_blocking=True)
            slot.d2h_done.record(self.s_d2h)
        slot.keep = (res, outs)
        slot.busy = True
        return i

    def wait(self, ticket: int) -> Dict[str, torch.Tensor]:
        slot = self.slots[ticket]
        slot.d2h_done.synchronize()
        slot.busy = False
        return slot.host_out

    def run(self, hosts):
        """Generator: results of every host batch in order, ``depth`` batches in flight."""
        pending: List[int] = []
        for host in hosts:
            if len(pending) == self.depth:
                yield self.wait(pending.pop(0))
            pending.append(self.submit(host))
        for t in pending:
            yield self.wait(t)

    # ------------------------------------------------------------------------------------------
    def d2h_bytes(self) -> int:
        out = self.slots[0].host_out or {}
        return int(sum(b.numel() * b.element_size() for b in out.values()))
