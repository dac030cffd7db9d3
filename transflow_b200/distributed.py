"""Frame-pair sharding across the GPUs of one box (one process per GPU).

Flow estimation of pair t is independent of pair t+-1 (Farneback with fb_flags = 0, PyrLK,
Horn-Schunck with hs_decay = 0), so it shards by CHUNKS of K consecutive pairs: a rank that
owns a chunk prepares K + 1 frames and reuses every frame's pyramid / polynomial expansion for
two pairs.  Accumulate + remap is a strict recurrence (``data_t = f(data_{t-1}, flow_t)``) and
runs on rank 0 only, in frame order.  There is no collective on the data path: the only
exchange is the point-to-point hand-off of finished flow fields to rank 0
(``torch.distributed`` batched send/recv = NCCL over NVLink on the GPU box, gloo in the CPU
tests).

The stream is cut into rounds.  In every round each producer rank r >= 1 owns ``q`` chunks and
rank 0 owns ``c0 <= q`` chunks, where ``c0`` balances rank 0's extra accumulate work (see
``plan_round``): with F = cost of one flow and A = cost of one accumulate+remap, rank 0 finishes
a round together with the producers when ``c0 = q * (1 - (N - 1) * A/F) / (1 + A/F)``.
Rank 0 posts the receives of round j + 1 before it consumes round j (two buffer sets), so
producers run one round ahead of the accumulator instead of stalling on it.
"""
import numpy as np
import torch
import torch.distributed as dist


def plan_round(world: int, q: int, flow_ms: float, accumulate_ms: float) -> list:
    """Chunks per rank in one round: ``[c0, q, q, ...]``.

    Rank 0 spends ``c0 * F + T * A`` per round (T = c0 + (world - 1) * q chunks), a producer
    ``q * F``.  Balancing gives ``c0 = q * (1 - (world - 1) * a) / (1 + a)`` with ``a = A / F``,
    clamped to ``[0, q]`` and rounded down (rank 0 must never be the late one).
    """
    if world == 1:
        return [q]
    a = accumulate_ms / max(flow_ms, 1e-9)
    c0 = q * (1.0 - (world - 1) * a) / (1.0 + a)
    c0 = int(max(0, min(q, np.floor(c0 + 1e-9))))
    return [c0] + [q] * (world - 1)


def round_slots(counts: list) -> list:
    """Owner rank of every chunk slot of a round, in stream order (rank 0's chunks first)."""
    owners = []
    for rank, c in enumerate(counts):
        owners += [rank] * c
    return owners


class ShardedFlowStream:
    """Drives one rank of the sharded pipeline.

    ``estimate_chunk(first_pair, n_pairs)`` -> list of post-processed flow tensors (producer side);
    ``accumulate(flow)`` consumes flows strictly in frame order (rank 0 only).
    Works with any ``torch.distributed`` backend: tensors only need to live on the device the
    backend moves (CUDA for nccl, CPU for gloo).
    """

    def __init__(self, rank, world, chunk_pairs, counts, estimate_chunk, accumulate, flow_shape, device,
                 group=None, transport="nccl", round_hook=None):
        self.rank, self.world, self.k = rank, world, chunk_pairs
        self.transport = transport
        self.round_hook = round_hook      # called on every rank with the round index before the round starts
        self.counts = list(counts)
        self.owners = round_slots(self.counts)
        self.cpr = len(self.owners)
        self.estimate_chunk, self.accumulate = estimate_chunk, accumulate
        self.flow_shape, self.device, self.group = tuple(flow_shape), device, group
        self.frames_accumulated = 0
        self._recv = {}      # (round parity, slot) -> K receive buffers
        self.ring = None
        if transport == "p2p" and world > 1:
            from .peer import PeerFlowRing
            flows_per_round = [c * chunk_pairs for c in self.counts]
            self.ring = PeerFlowRing(rank, world, flows_per_round, self.flow_shape, group)
            self._published = 0                       # producer: flows published so far
            self._expected = [0] * world              # rank 0: flows consumed so far per producer

    @property
    def frames_per_round(self) -> int:
        return self.cpr * self.k

    def _recv_buffers(self, parity: int, slot: int):
        key = (parity, slot)
        if key not in self._recv:
            self._recv[key] = [torch.empty(self.flow_shape, dtype=torch.float32, device=self.device)
                               for _ in range(self.k)]
        return self._recv[key]

    def _post_receives(self, round_index: int) -> dict:
        """One batched P2P group per remote chunk slot of the round -> {slot: [work, ...]}."""
        pending = {}
        for slot, owner in enumerate(self.owners):
            if owner != 0:
                bufs = self._recv_buffers(round_index & 1, slot)
                ops = [dist.P2POp(dist.irecv, b, owner, self.group) for b in bufs]
                pending[slot] = dist.batch_isend_irecv(ops)
        return pending

    def _run_p2p(self, first_round: int, n_rounds: int):
        """Same schedule; flows land in rank 0's ring by peer stores, counters replace send/recv."""
        ring = self.ring
        for j in range(first_round, first_round + n_rounds):
            base_chunk = j * self.cpr
            if self.round_hook is not None:
                self.round_hook(j)
            if self.rank == 0:
                index = [0] * self.world              # next ring slot per producer in this round
                for slot, owner in enumerate(self.owners):
                    if owner == 0:
                        flows = self.estimate_chunk((base_chunk + slot) * self.k, self.k)
                    else:
                        self._expected[owner] += self.k
                        ring.wait_ready(owner, self._expected[owner])
                        flows = [ring.slot_tensor(owner, j, index[owner] + i) for i in range(self.k)]
                        index[owner] += self.k
                    for f in flows:
                        self.accumulate(f)
                        self.frames_accumulated += 1
                    if owner != 0 and index[owner] == self.counts[owner] * self.k:
                        ring.release(owner, j)        # every flow of this producer's round is consumed
            else:
                ring.wait_slot_free(j)
                i = 0
                for slot, owner in enumerate(self.owners):
                    if owner == self.rank:
                        outs = [ring.slot_address(j, i + n) for n in range(self.k)]
                        self.estimate_chunk((base_chunk + slot) * self.k, self.k, outs)
                        i += self.k
                        self._published += self.k
                        ring.publish(self._published)

    def run(self, first_round: int, n_rounds: int):
        """Process rounds [first_round, first_round + n_rounds)."""
        if n_rounds <= 0:
            return
        if self.ring is not None:
            return self._run_p2p(first_round, n_rounds)
        last = first_round + n_rounds - 1
        if self.rank == 0:
            pending = self._post_receives(first_round)
            for j in range(first_round, last + 1):
                if self.round_hook is not None:
                    self.round_hook(j)
                upcoming = self._post_receives(j + 1) if j < last else {}
                base_chunk = j * self.cpr
                for slot, owner in enumerate(self.owners):
                    if owner == 0:
                        flows = self.estimate_chunk((base_chunk + slot) * self.k, self.k)
                    else:
                        for w in pending[slot]:
                            w.wait()
                        flows = self._recv_buffers(j & 1, slot)
                    for f in flows:
                        self.accumulate(f)
                        self.frames_accumulated += 1
                pending = upcoming
        else:
            in_flight = []       # (works, tensors) of the previous round: bounded look-ahead
            for j in range(first_round, last + 1):
                if self.round_hook is not None:
                    self.round_hook(j)
                base_chunk = j * self.cpr
                works, keep = [], []
                for slot, owner in enumerate(self.owners):
                    if owner == self.rank:
                        flows = [f.contiguous() for f in self.estimate_chunk((base_chunk + slot) * self.k, self.k)]
                        keep.append(flows)              # alive until the sends have completed
                        ops = [dist.P2POp(dist.isend, f, 0, self.group) for f in flows]
                        works += dist.batch_isend_irecv(ops)
                in_flight.append((works, keep))
                if len(in_flight) > 1:
                    for w in in_flight.pop(0)[0]:
                        w.wait()
            for works, _ in in_flight:
                for w in works:
                    w.wait()

    def run_round(self, round_index: int):
        self.run(round_index, 1)
