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


def plan_rank0_pairs(world: int, q: int, k: int, flow_ms: float, accumulate_ms: float) -> int:
    """Pairs rank 0 estimates itself per round when its share is counted in PAIRS, not whole chunks: with every
    producer owning ``q * k`` pairs, rank 0 (which also accumulates all ``p0 + (world - 1) * q * k`` frames) finishes
    with the producers for ``p0 = q k (F - (world - 1) A) / (F + A)``, rounded down, in ``[0, q k]``.  At 8 ranks the
    whole-chunk plan leaves rank 0 without any estimation work although a quarter of its time is free."""
    if world == 1:
        return q * k
    f, a = max(flow_ms, 1e-9), max(accumulate_ms, 0.0)
    p0 = q * k * (f - (world - 1) * a) / (f + a)
    return int(max(0, min(q * k, np.floor(p0 + 1e-9))))


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

    A round is a list of chunk slots ``(owner, pairs)`` in stream order.  By default rank 0's ``counts[0]`` chunks
    come first, then every producer's; with ``rank0_pairs`` given, rank 0 instead owns ONE chunk of that many pairs
    at the END of the round, and its estimation is queued ONE ROUND AHEAD (``estimate_chunk(..., join=False)`` on the
    estimator's own streams; the returned list may carry a ``ready()`` method that orders the caller's stream after
    that chunk only): like a producer, rank 0's estimator then runs a round ahead of the accumulation, which never
    waits for it.
    """

    def __init__(self, rank, world, chunk_pairs, counts, estimate_chunk, accumulate, flow_shape, device,
                 group=None, transport="nccl", round_hook=None, rank0_pairs=None):
        self.rank, self.world, self.k = rank, world, chunk_pairs
        self.transport = transport
        self.round_hook = round_hook      # called on every rank with the round index before the round starts
        self.counts = list(counts)
        self.rank0_pairs = rank0_pairs
        if rank0_pairs is None:
            self.slots = [(owner, chunk_pairs) for owner in round_slots(self.counts)]
        else:
            self.slots = [(owner, chunk_pairs) for owner in round_slots([0] + self.counts[1:])]
            if rank0_pairs > 0:
                self.slots.append((0, int(rank0_pairs)))
        self.owners = [owner for owner, _ in self.slots]
        self.offsets = [0]
        for _, n in self.slots:
            self.offsets.append(self.offsets[-1] + n)
        self.cpr = len(self.slots)
        self.estimate_chunk, self.accumulate = estimate_chunk, accumulate
        self.flow_shape, self.device, self.group = tuple(flow_shape), device, group
        self.frames_accumulated = 0
        self._ahead = {}     # round -> {slot: flows} of rank 0's own chunks queued ahead (rank0_pairs mode)
        self._recv = {}      # (round parity, slot) -> receive buffers
        self.ring = None
        if transport == "p2p" and world > 1:
            from .peer import PeerFlowRing
            flows_per_round = [sum(n for owner, n in self.slots if owner == r) for r in range(world)]
            self.ring = PeerFlowRing(rank, world, flows_per_round, self.flow_shape, group)
            self._published = 0                       # producer: flows published so far
            self._expected = [0] * world              # rank 0: flows consumed so far per producer

    @property
    def frames_per_round(self) -> int:
        return self.offsets[-1]

    def _first_pair(self, round_index: int, slot: int) -> int:
        return round_index * self.frames_per_round + self.offsets[slot]

    def _recv_buffers(self, parity: int, slot: int):
        key = (parity, slot)
        if key not in self._recv:
            self._recv[key] = [torch.empty(self.flow_shape, dtype=torch.float32, device=self.device)
                               for _ in range(self.slots[slot][1])]
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

    def _own_chunks_ahead(self, round_index: int, last_round: int) -> dict:
        """rank0_pairs mode: make sure rank 0's own chunks of this round AND the next one (inside the run) are queued,
        without ordering the caller's stream after them; returns this round's."""
        if self.rank0_pairs is None:
            return {}
        for jj in (round_index, round_index + 1):
            if jj <= last_round and jj not in self._ahead:
                self._ahead[jj] = {slot: self.estimate_chunk(self._first_pair(jj, slot), n, join=False)
                                   for slot, (owner, n) in enumerate(self.slots) if owner == 0}
        return self._ahead.pop(round_index)

    @staticmethod
    def _own_flows(flows):
        if hasattr(flows, "ready"):
            flows.ready()         # the caller's stream waits for that chunk (not for chunks queued after it)
        return flows

    def _run_p2p(self, first_round: int, n_rounds: int):
        """Same schedule; flows land in rank 0's ring by peer stores, counters replace send/recv."""
        ring = self.ring
        last_round = first_round + n_rounds - 1
        for j in range(first_round, first_round + n_rounds):
            if self.round_hook is not None:
                self.round_hook(j)
            if self.rank == 0:
                ahead = self._own_chunks_ahead(j, last_round)
                index = [0] * self.world              # next ring slot per producer in this round
                total = [sum(n for o, n in self.slots if o == r) for r in range(self.world)]
                for slot, (owner, n) in enumerate(self.slots):
                    if owner == 0:
                        if slot in ahead:
                            flows = self._own_flows(ahead[slot])
                        else:
                            flows = self.estimate_chunk(self._first_pair(j, slot), n)
                    else:
                        self._expected[owner] += n
                        ring.wait_ready(owner, self._expected[owner])
                        flows = [ring.slot_tensor(owner, j, index[owner] + i) for i in range(n)]
                        index[owner] += n
                    for f in flows:
                        self.accumulate(f)
                        self.frames_accumulated += 1
                    if owner != 0 and index[owner] == total[owner]:
                        ring.release(owner, j)        # every flow of this producer's round is consumed
            else:
                i = 0
                first = True
                for slot, (owner, n) in enumerate(self.slots):
                    if owner == self.rank:
                        outs = [ring.slot_address(j, i + m) for m in range(n)]
                        # the ring's "slot free" wait of a new round goes on every stream the estimator stores from
                        gate = (lambda jj=j: ring.wait_slot_free(jj)) if first else None
                        self.estimate_chunk(self._first_pair(j, slot), n, outs, gate)
                        first = False
                        i += n
                        self._published += n
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
                ahead = self._own_chunks_ahead(j, last)
                for slot, (owner, n) in enumerate(self.slots):
                    if owner == 0:
                        if slot in ahead:
                            flows = self._own_flows(ahead[slot])
                        else:
                            flows = self.estimate_chunk(self._first_pair(j, slot), n)
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
                works, keep = [], []
                for slot, (owner, n) in enumerate(self.slots):
                    if owner == self.rank:
                        flows = [f.contiguous() for f in self.estimate_chunk(self._first_pair(j, slot), n)]
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
