"""Host-facing steady-state loop of the pair stage: pinned host batches in, pinned host results out.

The reference's driver (lib/modeling/predict.py:42-57) feeds one ``PairList`` at a time from a
``DataLoader`` and reads the outputs back with ``.numpy()``.  On a B200 a per-video call is
launch- and PCIe-latency-bound, so the serving loop works on *batches of videos* and keeps
``depth`` of them in flight over three streams:

    h2d stream    :  H2D(i+1) ......................
    compute stream:  kernels(i)  (one CUDA-graph launch, see pipeline.GraphedStage); two compute streams
                     alternate, so the end of step i overlaps the start of step i+1
    d2h stream    :  D2H(i-1) ......................

Batches are ragged in practice (every video has its own tracklet and frame count).  The loop therefore
owns one *bucket* per ``batch.Capacity`` (normally one per chunk class of the pair kernel, see
``batch.pack_batches``): ``depth`` slots, each with its device input buffers, its captured graph (and
therefore its output buffers) and its pinned host result buffers.  Any batch packed for a bucket's
capacity replays that bucket's graph - no re-capture, nothing allocated in steady state; a batch
without a capacity gets a bucket for its exact per-video shapes.  With ``group`` the per-video triplet
records of each step are all-gathered across ranks (the one collective of the path, sharding.py) on a
stream of their own: only the D2H copy of the gathered records waits for it (every rank must then
submit the same sequence of buckets).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch

from .batch import Capacity, DeviceBatch, HostBatch, bucket_capacities, pack_batches
from .pipeline import GraphedStage, PairStage


def bucket_key(host: HostBatch):
    if host.capacity is not None:
        return host.capacity
    return ("exact", tuple(host.n), tuple(host.t), host.boxes_compact, host.motion_compact,
            tuple(sorted(k for k in host.layout if k != "bytes")))


class _Slot:
    def __init__(self, stage: PairStage, template: HostBatch, device, graphs: bool, single_graph: bool = True):
        self.batch: DeviceBatch = template.to_device(device, non_blocking=False)
        torch.cuda.synchronize(device)
        self.graphed: Optional[GraphedStage] = stage.capture(self.batch, single=single_graph) if graphs else None
        self.h2d_done = torch.cuda.Event()
        self.kernels_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.pinned: Optional[Dict[str, torch.Tensor]] = None      # capacity-sized pinned result buffers
        self.host_out: Optional[Dict[str, torch.Tensor]] = None    # views of them for the batch in flight
        self.gathered: Optional[torch.Tensor] = None
        self.keep = None
        self.busy = False


class _Bucket:
    def __init__(self, slots: List[_Slot]):
        self.slots = slots
        self.next = 0


class PipelinedStage:
    """``submit(host_batch)`` enqueues H2D -> kernels -> D2H for one batch and returns a ticket;
    ``wait(ticket)`` blocks until that batch's results are in pinned host memory and returns them
    (valid until ``depth`` further submits to the same bucket).  ``templates``: one host batch per bucket
    (a single batch is accepted too); further buckets are created on first use (which captures a graph:
    milliseconds - pass every capacity up front to keep the steady state capture-free)."""

    def __init__(self, stage: PairStage, templates, device="cuda", depth: int = 2, graphs: bool = True,
                 group=None, compute_streams: int = 2, single_graph: bool = True, collective: str = "nccl"):
        """``collective`` (with ``group``): ``"nccl"`` = one ``all_gather_into_tensor`` of the records per step on a
        stream of its own; ``"peer"`` = every rank stores its records into every rank's gather buffer over NVLink
        (``sharding.PeerRecords``: no collective launch; needs peer-mappable memory, one bucket)."""
        self.stage, self.device, self.depth, self.group = stage, torch.device(device), int(depth), group
        self.collective = collective if group is not None else None
        self.peer = None
        self.graphs, self.single_graph = bool(graphs), bool(single_graph)
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
        self.buckets: Dict[object, _Bucket] = {}
        self._submits = 0
        self._d2h_bytes = 0
        if isinstance(templates, HostBatch):
            templates = [templates]
        for t in templates:
            self._bucket(t)

    def _bucket(self, host: HostBatch) -> _Bucket:
        key = bucket_key(host)
        b = self.buckets.get(key)
        if b is None:
            b = _Bucket([_Slot(self.stage, host, self.device, self.graphs, self.single_graph)
                         for _ in range(self.depth)])
            self.buckets[key] = b
        return b

    @property
    def slots(self) -> List[_Slot]:
        """Slots of the first bucket (the only one of a fixed-shape loop)."""
        return next(iter(self.buckets.values())).slots

    # ------------------------------------------------------------------------------------------
    def submit(self, host: HostBatch, post=None):
        """``post(result)``: optional callable run on the step's compute stream right behind its kernels (e.g. to
        stash the records of a sharded run for the collective at the end of a pass)."""
        bucket = self._bucket(host)
        i = bucket.next
        bucket.next = (i + 1) % self.depth
        slot = bucket.slots[i]
        if slot.busy:
            raise RuntimeError("PipelinedStage: %d batches of this bucket already in flight; wait() for a ticket first"
                               % self.depth)
        with torch.cuda.stream(self.s_h2d):
            self.s_h2d.wait_event(slot.kernels_done)     # the slot's previous kernels have consumed its inputs
            slot.batch.copy_from(host)
            slot.h2d_done.record(self.s_h2d)
        main = self.compute[self._submits % len(self.compute)]
        self._submits += 1
        with torch.cuda.stream(main):
            main.wait_event(slot.h2d_done)
            main.wait_event(slot.d2h_done)               # the slot's previous results have left the device
            res = slot.graphed.replay() if slot.graphed is not None else self.stage.forward(slot.batch)
            outs = res.host_outputs()
            if post is not None:
                post(res)
            slot.kernels_done.record(main)
        work = None
        if self.group is not None and self.collective == "peer":
            if self.peer is None:                        # first step: symmetric buffers + rendezvous (collective call)
                from .sharding import PeerRecords
                with torch.cuda.stream(main):
                    self.peer = PeerRecords(res.records.shape, self.group, self.device, ring=self.depth + 2)
            with torch.cuda.stream(main):                # behind the step's kernels: plain stores into the peers
                self.peer.scatter(res.records)
                slot.kernels_done.record(main)
        elif self.group is not None:                     # the one collective: top-K triplet records
            import torch.distributed as dist
            world = dist.get_world_size(self.group)
            if slot.gathered is None:
                slot.gathered = torch.empty((world,) + tuple(res.records.shape), dtype=res.records.dtype,
                                            device=self.device)
            with torch.cuda.stream(self.s_comm):
                self.s_comm.wait_event(slot.kernels_done)
                work = dist.all_gather_into_tensor(slot.gathered, res.records, group=self.group, async_op=True)
            outs["records_all_ranks"] = slot.gathered
        peer_view = None
        if self.peer is not None:
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(slot.kernels_done)
                peer_view = self.peer.gather()           # waits (on this stream only) for every rank's step
            outs["records_all_ranks"] = peer_view
        if slot.graphed is None or slot.pinned is None:
            # first use (or eager mode, whose outputs are fresh tensors): pinned buffers of the bucket's full size
            full = res.host_outputs(full=True)
            if self.group is not None:
                full["records_all_ranks"] = peer_view if peer_view is not None else slot.gathered
            if slot.pinned is None or any(k not in slot.pinned or slot.pinned[k].shape != v.shape
                                          for k, v in full.items()):
                slot.pinned = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in full.items()}
        slot.host_out = {}
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(slot.kernels_done)
            if work is not None:
                work.wait()                              # this stream waits for the all-gather, nobody else does
            for k, src in outs.items():
                if slot.graphed is None:
                    src.record_stream(self.s_d2h)
                dst = slot.pinned[k][:src.shape[0]]
                dst.copy_(src, non_blocking=True)
                slot.host_out[k] = dst
                self._d2h_bytes += src.numel() * src.element_size()
            if peer_view is not None:
                self.peer.release()                      # the slot of the gather buffer may be overwritten again
            slot.d2h_done.record(self.s_d2h)
        slot.keep = (res, outs)
        slot.busy = True
        return (bucket, i)

    def wait(self, ticket) -> Dict[str, torch.Tensor]:
        bucket, i = ticket
        slot = bucket.slots[i]
        slot.d2h_done.synchronize()
        slot.busy = False
        return slot.host_out

    def run(self, hosts: Iterable[HostBatch], post=None):
        """Generator: results of every host batch in order, up to ``depth`` batches in flight.  ``post(q, result)``
        is called for the q-th batch on its compute stream right behind its kernels (see ``submit``)."""
        pending: List = []
        for q, host in enumerate(hosts):
            b = self._bucket(host)
            # at most `depth` batches in flight, and the slot this batch will take must be free
            while pending and (len(pending) >= self.depth or b.slots[b.next].busy):
                yield self.wait(pending.pop(0))
            pending.append(self.submit(host, post=(lambda res, q=q: post(q, res)) if post is not None else None))
        for t in pending:
            yield self.wait(t)

    # ------------------------------------------------------------------------------------------
    def d2h_total_bytes(self) -> int:
        """Bytes copied device -> host by every batch submitted so far."""
        return int(self._d2h_bytes)

    def d2h_bytes(self) -> int:
        """Bytes the last batch submitted to the first bucket's first used slot copied device -> host."""
        for b in self.buckets.values():
            for s in b.slots:
                if s.host_out:
                    return int(sum(t.numel() * t.element_size() for t in s.host_out.values()))
        return 0


def host_batches_for(videos: Sequence, n_classes: int, geo_budget_bytes: int = 4 << 30, max_videos: int = 64,
                     capacities: Optional[Dict[int, Capacity]] = None, pin: bool = True,
                     merge_below_bytes: int = 512 << 20):
    """Ragged videos -> ``(host batches, [video indices of each batch], {t_class: Capacity})``: the videos are packed
    into batches per chunk class (``batch.pack_batches``), one capacity per class holds all of them."""
    shapes = [(int(v.boxes.shape[0]), int(v.boxes.shape[1])) for v in videos]
    batches = pack_batches(shapes, geo_budget_bytes, max_videos, merge_below_bytes)
    caps = capacities or bucket_capacities(shapes, batches, n_classes)
    hosts = [HostBatch.from_videos([videos[i] for i in vids], pin=pin, capacity=caps[c]) for c, vids in batches]
    return hosts, [vids for _, vids in batches], caps
