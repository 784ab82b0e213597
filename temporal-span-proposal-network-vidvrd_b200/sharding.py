"""Per-video sharding across GPUs and the one collective of the pair stage.

Videos are independent units (the reference treats every ``(vid, fstart, fend)`` as its own dataset
item, lib/dataset/vrdataset.py:56-83, and regroups per video at base.py:92-96), so the path shards
with no data-path collective: each rank runs the whole pair stage on its own videos.  The only
exchange is the all-gather of the fixed-size top-K triplet records at the end (32 bytes x
TOPK_PER_SEG per video), which replaces the pickle-based ``comm.all_gather`` of
lib/utils/comm.py:48-88 (two all-gathers plus padding) by a single ``all_gather_into_tensor``.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def video_cost(n_tracklets: int, n_frames: int) -> int:
    """Work of one video: ordered pairs x frames (the geometry kernel's bytes are 32*T per pair)."""
    return int(n_tracklets) * max(int(n_tracklets) - 1, 0) * int(n_frames)


def shard_videos(shapes: Sequence[Tuple[int, int]], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of videos to ranks.

    ``shapes[i] = (N_i, T_i)``.  Returns ``world_size`` lists of video indices (each sorted
    ascending); deterministic: ties go to the lower video index, then to the lower rank.  The
    reference's ``DistributedSampler`` (lib/dataset/samplers/distributed.py:49-58) is the naive
    equivalent: a contiguous split with wrap-around padding and no cost model.
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(shapes)), key=lambda i: (-video_cost(*shapes[i]), i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        out[r].append(i)
        loads[r] += video_cost(*shapes[i])
    return [sorted(v) for v in out]


def imbalance(shapes: Sequence[Tuple[int, int]], shards: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(video_cost(*shapes[i]) for i in s) for s in shards]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0


def gather_records(records: torch.Tensor, counts: torch.Tensor, shards: List[List[int]], group=None):
    """All-gather the per-video triplet records of every rank and put them back in video order.

    ``records [V_local, M, 8]`` int32, ``counts [V_local]`` int32 for this rank's videos (in the order
    of ``shards[rank]``).  Ranks hold different numbers of videos, so each pads to ``max_v`` videos;
    one ``all_gather_into_tensor`` moves ``world * max_v * (M*8 + 1)`` int32.  Returns
    ``(records_all [V_total, M, 8], counts_all [V_total])`` indexed by global video id, identical on
    every rank.  Works on NCCL (CUDA tensors) and gloo (CPU tensors).
    """
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    v_local, m = int(records.shape[0]), int(records.shape[1])
    if v_local != len(shards[rank]):
        raise ValueError("rank %d holds %d videos but its shard lists %d" % (rank, v_local, len(shards[rank])))
    max_v = max(len(s) for s in shards)
    width = m * 8 + 1
    send = torch.zeros((max_v, width), dtype=torch.int32, device=records.device)
    if v_local:
        send[:v_local, :m * 8] = records.reshape(v_local, m * 8)
        send[:v_local, m * 8] = counts
    recv = torch.empty((world, max_v, width), dtype=torch.int32, device=records.device)
    dist.all_gather_into_tensor(recv.view(world * max_v, width), send, group=group)
    total = sum(len(s) for s in shards)
    out_r = torch.zeros((total, m, 8), dtype=torch.int32, device=records.device)
    out_c = torch.zeros(total, dtype=torch.int32, device=records.device)
    for r, vids in enumerate(shards):
        if vids:
            idx = torch.as_tensor(vids, dtype=torch.int64, device=records.device)
            out_r[idx] = recv[r, :len(vids), :m * 8].reshape(len(vids), m, 8)
            out_c[idx] = recv[r, :len(vids), m * 8]
    return out_r, out_c


class PeerRecords:
    """The per-step exchange of the triplet records over peer memory (``csrc/peer_records.cu``): every rank stores its
    ``[V, M, 8]`` records straight into every rank's gather buffer over NVLink and publishes a flag; no collective is
    launched.  Buffers are ``torch.distributed._symmetric_memory`` allocations (peer-mapped by ``rendezvous``).

        peer = PeerRecords(records_shape, group, device, ring=depth + 2)
        peer.scatter(records)          # producer, on the stream that produced `records`
        gathered = peer.gather()       # consumer, on the stream that reads them out: waits for every rank's step,
        ...copy `gathered` out...      # returns this step's [world, V, M, 8] view
        peer.release()                 # consumer, after the copy: the slot may be overwritten ``ring`` steps later

    Every rank must run the same sequence of steps.  ``check()`` raises if a bounded wait on the device expired.
    ``PeerRecords.available(group)`` tells whether the ranks of ``group`` can map each other's memory; the caller
    falls back to ``all_gather_into_tensor`` otherwise."""

    MAX_SPIN = 5_000_000            # x ~0.2 us: a second before a wait gives up

    def __init__(self, records_shape, group, device, ring: int = 5):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self._lib = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.ring = int(ring)
        self.shape = tuple(int(x) for x in records_shape)
        self.n_int32 = 1
        for x in self.shape:
            self.n_int32 *= x
        if self.n_int32 % 4:
            raise ValueError("records must be a multiple of 16 bytes")
        dev = torch.device(device)
        self.buf = symm.empty((self.ring, self.world, self.n_int32), dtype=torch.int32, device=dev)
        self.flags = symm.empty((2, self.ring, self.world), dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.flags.zero_()
        h_buf = symm.rendezvous(self.buf, self.group)
        h_flg = symm.rendezvous(self.flags, self.group)
        bufs = [h_buf.get_buffer(r, tuple(self.buf.shape), torch.int32).data_ptr() for r in range(self.world)]
        flgs = [h_flg.get_buffer(r, tuple(self.flags.shape), torch.int32).data_ptr() for r in range(self.world)]
        self.peer_bufs = torch.tensor(bufs, dtype=torch.int64, device=dev)
        self.peer_flags = torch.tensor(flgs, dtype=torch.int64, device=dev)
        self.state = torch.zeros(2 + self.ring, dtype=torch.int32, device=dev)
        self._h = (h_buf, h_flg)
        self.produced = self.consumed = 0           # exchanges issued so far (host counters: the kernels' step numbers)
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)              # every rank's flags are zero before anyone publishes

    @staticmethod
    def available(group=None) -> bool:
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory  # noqa: F401
            return dist.is_initialized() and dist.get_backend(group) == "nccl" and torch.cuda.is_available()
        except Exception:  # noqa: BLE001
            return False

    def _stream(self):
        return torch.cuda.current_stream(self.buf.device).cuda_stream

    def scatter(self, records: torch.Tensor) -> None:
        records = records.contiguous()
        if records.numel() != self.n_int32 or records.dtype != torch.int32:
            raise ValueError("records %s do not match the exchange's shape %s" % (tuple(records.shape), self.shape))
        if self.produced - self.consumed >= self.ring - 1:
            raise RuntimeError("peer record exchange: %d steps in flight on a ring of %d" % (self.produced - self.consumed,
                                                                                          self.ring))
        self.produced += 1
        self._lib.check(self._lib.load().tspn_records_scatter(
            records.data_ptr(), self.n_int32, self.peer_bufs.data_ptr(), self.peer_flags.data_ptr(),
            self.flags.data_ptr(), self.world, self.rank, self.ring, self.produced, self.state.data_ptr(), self.MAX_SPIN,
            self._stream()), "tspn_records_scatter")

    def gather(self) -> torch.Tensor:
        """Wait (on the current stream) for every rank's records of the next step; returns their ``[world, *shape]``
        view in the gather buffer - valid until ``release()`` and ``ring - 1`` further steps."""
        self.consumed += 1
        self._lib.check(self._lib.load().tspn_records_wait(self.flags.data_ptr(), self.world, self.ring, self.consumed,
                                                           self.state.data_ptr(), self.MAX_SPIN, self._stream()),
                        "tspn_records_wait")
        return self.buf[self.consumed % self.ring].view((self.world,) + self.shape)

    def release(self) -> None:
        """The step returned by the last ``gather()`` has been read (on the current stream): its slot may be reused."""
        self._lib.check(self._lib.load().tspn_records_release(self.peer_flags.data_ptr(), self.world, self.rank,
                                                              self.ring, self.consumed, self._stream()),
                        "tspn_records_release")

    def check(self) -> None:
        err = int(self.state[0].item())
        if err:
            raise RuntimeError("peer record exchange: a bounded wait expired on the device (%s)"
                               % ("credits of a slot" if err == 1 else "records of a step"))
