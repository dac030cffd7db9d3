"""Direct flow hand-off over NVLink peer memory (transport "p2p" of ``distributed.py``).

Rank 0 owns, per producer rank, a ring of flow slots (two round parities) and a ``ready``
counter; every producer owns a ``consumed`` counter.  The buffers are raw ``cudaMalloc``
allocations exchanged as CUDA-IPC handles.  A producer's LAST kernel for a pair (the
post-process gather / clip) stores the finished flow straight into rank 0's slot, then a
one-thread kernel publishes the producer's running flow count with a system-scope release
store; rank 0's stream waits for the count with ``cuStreamWaitValue32`` -- no kernel ever spins
on memory another rank writes.  Back-pressure runs the same way in the other direction.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, stream_ptr


class _RawArray:
    def __init__(self, address: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 2}


class DeviceBuffer:
    """A raw, zero-initialised device allocation that can be shared through CUDA IPC."""

    def __init__(self, nbytes: int):
        self.lib = _lib.load()
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(self.lib.tf_device_malloc(self.nbytes, C.byref(p)))
        self.address = int(p.value)

    def handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        check(self.lib.tf_ipc_get_handle(C.c_void_p(self.address), buf))
        return bytes(buf)

    def tensor(self, offset: int, shape, dtype=torch.float32) -> torch.Tensor:
        n = int(torch.tensor([], dtype=dtype).element_size())
        for s in shape:
            n *= int(s)
        raw = torch.as_tensor(_RawArray(self.address + offset, n), device="cuda")
        return raw.view(dtype).view(*shape)

    def free(self):
        if self.address:
            self.lib.tf_device_free(C.c_void_p(self.address))
            self.address = 0


def raw_tensor(address: int, shape, dtype=torch.float32) -> torch.Tensor:
    """A tensor view of raw device memory at ``address`` (local, or peer memory mapped through CUDA IPC): what
    ``Tensor.copy_`` needs to move a buffer with the copy engines instead of a kernel."""
    n = int(torch.tensor([], dtype=dtype).element_size())
    for s in shape:
        n *= int(s)
    return torch.as_tensor(_RawArray(int(address), n), device="cuda").view(dtype).view(*shape)


def open_ipc(handle: bytes) -> int:
    p = C.c_void_p()
    buf = (C.c_uint8 * 64).from_buffer_copy(handle)
    check(_lib.load().tf_ipc_open_handle(buf, C.byref(p)))
    return int(p.value)


def flag_signal(address: int, value: int):
    check(_lib.load().tf_flag_signal(C.c_void_p(address), C.c_uint32(value), stream_ptr()))


def flag_wait_geq(address: int, value: int):
    check(_lib.load().tf_flag_wait_geq(C.c_void_p(address), C.c_uint32(value), stream_ptr()))


class PeerFlowRing:
    """Ring of flow slots on rank 0, one lane per producer; see the module docstring."""

    def __init__(self, rank: int, world: int, slots_per_round: list, flow_shape, group=None):
        self.rank, self.world = rank, world
        self.slots = list(slots_per_round)             # flows per round for every rank (index 0 unused)
        self.flow_shape = tuple(flow_shape)
        self.flow_bytes = 4
        for s in self.flow_shape:
            self.flow_bytes *= int(s)
        self.local = {}
        mine = {}
        if rank == 0:
            for r in range(1, world):
                if self.slots[r] > 0:
                    ring = DeviceBuffer(2 * self.slots[r] * self.flow_bytes)
                    ready = DeviceBuffer(256)
                    self.local[r] = (ring, ready)
                    mine[f"ring{r}"] = ring.handle()
                    mine[f"ready{r}"] = ready.handle()
        else:
            self.consumed = DeviceBuffer(256)
            mine["consumed"] = self.consumed.handle()
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        if rank == 0:
            self.consumed_peer = {r: open_ipc(everyone[r]["consumed"]) for r in range(1, world) if self.slots[r] > 0}
        else:
            self.ring_peer = open_ipc(everyone[0][f"ring{rank}"]) if self.slots[rank] > 0 else 0
            self.ready_peer = open_ipc(everyone[0][f"ready{rank}"]) if self.slots[rank] > 0 else 0
        torch.cuda.synchronize()
        dist.barrier(group=group)

    # producer side -----------------------------------------------------------------------------
    def slot_address(self, round_index: int, i: int) -> int:
        return self.ring_peer + ((round_index & 1) * self.slots[self.rank] + i) * self.flow_bytes

    def wait_slot_free(self, round_index: int):
        if round_index >= 2:
            flag_wait_geq(self.consumed.address, round_index - 1)

    def publish(self, produced_total: int):
        flag_signal(self.ready_peer, produced_total)

    # rank 0 side ----------------------------------------------------------------------------------
    def wait_ready(self, producer: int, produced_total: int):
        flag_wait_geq(self.local[producer][1].address, produced_total)

    def slot_tensor(self, producer: int, round_index: int, i: int) -> torch.Tensor:
        ring = self.local[producer][0]
        return ring.tensor(((round_index & 1) * self.slots[producer] + i) * self.flow_bytes, self.flow_shape)

    def release(self, producer: int, round_index: int):
        flag_signal(self.consumed_peer[producer], round_index + 1)


class PeerFrameFanout:
    """Spreads rank 0's finished RGB frames over every rank's PCIe link.

    The accumulate+remap recurrence runs on rank 0 only, so at N GPUs its single PCIe link would
    have to carry every output frame (25 MB at 4K).  Instead rank 0 stores frame i into the output
    ring of rank i % N over NVLink (a kernel doing 128-bit peer stores, then a release flag), and
    that rank's copy stream -- waiting on the flag with a stream memory op -- moves it to its own
    pinned host buffer.  Per-slot ``freed`` counters flow back the same way.
    """

    def __init__(self, rank: int, world: int, frame_shape, slots: int = 4, group=None):
        self.rank, self.world, self.slots = rank, world, slots
        self.frame_shape = tuple(frame_shape)
        self.frame_bytes = 1
        for s in self.frame_shape:
            self.frame_bytes *= int(s)
        self.frame_bytes_padded = (self.frame_bytes + 255) // 256 * 256
        self.lib = _lib.load()
        mine = {}
        if rank == 0:
            self.freed = {r: DeviceBuffer(256) for r in range(1, world)}
            for r, buf in self.freed.items():
                mine[f"freed{r}"] = buf.handle()
            self.sent = [0] * world
        else:
            self.ring = DeviceBuffer(slots * self.frame_bytes_padded)
            self.filled = DeviceBuffer(256)
            mine["ring"] = self.ring.handle()
            mine["filled"] = self.filled.handle()
            self.received = 0
            self.stream = torch.cuda.Stream()
            self.host = [torch.empty(self.frame_shape, dtype=torch.uint8).pin_memory() for _ in range(slots)]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        if rank == 0:
            self.ring_peer = {r: open_ipc(everyone[r]["ring"]) for r in range(1, world)}
            self.filled_peer = {r: open_ipc(everyone[r]["filled"]) for r in range(1, world)}
        else:
            self.freed_peer = open_ipc(everyone[0][f"freed{rank}"])
        torch.cuda.synchronize()
        dist.barrier(group=group)

    # rank 0, on the compute stream, right after the frame was produced ----------------------------
    def send(self, frame: torch.Tensor, target: int):
        seq = self.sent[target] + 1
        slot = (seq - 1) % self.slots
        if seq > self.slots:
            flag_wait_geq(self.freed[target].address, seq - self.slots)
        dst = self.ring_peer[target] + slot * self.frame_bytes_padded
        check(self.lib.tf_copy_to_peer(C.c_void_p(dst), C.c_void_p(frame.data_ptr()), self.frame_bytes, stream_ptr()))
        flag_signal(self.filled_peer[target], seq)
        self.sent[target] = seq

    # receiving rank: enqueue the wait + D2H + release for the next n frames (does not block the host) ----
    def expect(self, n: int):
        with torch.cuda.stream(self.stream):
            for _ in range(n):
                seq = self.received + 1
                slot = (seq - 1) % self.slots
                flag_wait_geq(self.filled.address, seq)
                src = self.ring.tensor(slot * self.frame_bytes_padded, self.frame_shape, torch.uint8)
                self.host[slot].copy_(src, non_blocking=True)
                flag_signal(self.freed_peer, seq)
                self.received = seq
