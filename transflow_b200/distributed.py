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
import os

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


class _HostFrameFeeder:
    """Pinned host clip -> device frames through a side stream with one frame of look-ahead."""

    def __init__(self, frames_pinned, order_fn):
        self.frames, self.order = frames_pinned, order_fn
        self.stream = torch.cuda.Stream()
        self.cache = {}

    def _start(self, idx):
        if idx in self.cache:
            return
        with torch.cuda.stream(self.stream):
            dev = self.frames[self.order(idx)].cuda(non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.cache[idx] = (dev, ev)

    def get(self, idx):
        self._start(idx)
        dev, ev = self.cache.pop(idx)
        self._start(idx + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        dev.record_stream(cur)
        for stale in [k for k in self.cache if k != idx + 1]:
            del self.cache[stale]
        return dev


# ------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): the C3 workload sharded over N ranks
# ------------------------------------------------------------------------------------------------
def bench_sharded(args, rank, world, local):
    import json

    from . import _lib, ops
    from .compositor import Compositor
    from .compositor.pixmap_source_interface import PixmapSourceInterface, StillQueue
    from .config import LayerConfig
    import bench as B

    H, W = args.height, args.width
    K, Q = 8, 4
    clip, mask, pixmap = B.build_workload(H, W, B.N_DISTINCT, seed=0)
    frames_dev = torch.from_numpy(clip).cuda()
    feeder = _HostFrameFeeder(torch.from_numpy(clip).pin_memory(), lambda i: B.frame_order(i, B.N_DISTINCT))
    io = {"host": False}

    def frame(idx):
        return feeder.get(idx) if io["host"] else frames_dev[B.frame_order(idx, B.N_DISTINCT)]
    fb = ops.Farneback(H, W)
    # two pairs of a chunk in flight (handle lanes, tf_farneback_step_lane): pair i on lane i % 2 / its own stream,
    # frame i in slot i % 3; each lane has its own post-process scratch.  Chunks stay ordered on the caller's stream.
    lanes = max(1, min(2, int(os.environ.get("TFB200_FB_LANES", "2"))))
    nslots = 3 if lanes > 1 else 2
    posts = [ops.PostProcess(H, W, forward=True) for _ in range(lanes)]
    grays = [torch.empty((H, W), dtype=torch.uint8, device="cuda") for _ in range(lanes)]
    lane_streams = [torch.cuda.Stream() for _ in range(lanes)] if lanes > 1 else None

    chunk_state = {"next": None, "n": 0}

    def estimate_chunk(first_pair, n_pairs, outs=None):
        """K consecutive pairs: K solves + post-processes, and one extra prepare when the chunk does not continue
        the previous one (a rank's chunks of a round are consecutive in frame order: the lanes then run on without
        a new prepare and without draining).  ``outs`` (raw device addresses, possibly peer memory) receive the
        post-processed flows."""
        main = torch.cuda.current_stream()
        cont = lane_streams is not None and chunk_state["next"] == first_pair
        if not cont:
            ops.gray_from_bgr(frame(first_pair), grays[0])
            fb.prepare(0, grays[0])
            chunk_state["n"] = 0
            if lane_streams:        # also orders the lanes after the ring's "slot free" wait of a new round
                for s in lane_streams:
                    s.wait_stream(main)
        flows = []
        for i in range(n_pairs):
            n = chunk_state["n"]
            lane = n % lanes
            with torch.cuda.stream(lane_streams[lane] if lane_streams else main):
                old, new = n % nslots, (n + 1) % nslots
                ops.gray_from_bgr(frame(first_pair + i + 1), grays[lane])
                flow = fb.step(new, grays[lane], old, new, lane=lane)   # prepare(new) overlapped with solve: forward
                flow.record_stream(main)
                flows.append(posts[lane](flow, None if outs is None else outs[i]))
            chunk_state["n"] = n + 1
        chunk_state["next"] = first_pair + n_pairs
        if lane_streams:
            for s in lane_streams:
                main.wait_stream(s)
        return flows

    comp = None
    if rank == 0:
        mask_png = B.write_mask_png(mask, "shard")
        comp = Compositor.from_args(H, W, [LayerConfig(0, "moveref", reset_mode="random", reset_random_factor=0.5,
                                                       reset_mask=mask_png)], background_color=B.BG)
        comp.set_sources({0: [PixmapSourceInterface(StillQueue(torch.from_numpy(pixmap).cuda()),
                                                    np.ones((H, W), bool))]})
        rgb = [torch.empty((H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
        rgb_host = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        down = torch.cuda.Stream()
        copied = [None, None]
        count = {"n": 0}

    fanout_box = {"f": None, "frames": 0}

    def accumulate(flow):
        k = count["n"] & 1
        count["n"] += 1
        if io["host"] and copied[k] is not None:
            torch.cuda.current_stream().wait_event(copied[k])   # the D2H out of this buffer is done
        comp.step(flow, rgb[k])
        fan = fanout_box["f"]
        if io["host"] and fan is not None:
            target = fanout_box["frames"] % world            # frame i leaves through rank i % N's PCIe link
            fanout_box["frames"] += 1
            if target != 0:
                fan.send(rgb[k], target)
                done = torch.cuda.Event()
                done.record()
                copied[k] = done                              # rgb[k] may be rewritten once the peer copy ran
                return
        if io["host"]:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(down):
                down.wait_event(ready)
                rgb_host[k].copy_(rgb[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(down)
            copied[k] = ev

    # calibrate F (flow per pair) and A (accumulate per frame) on rank 0, share the plan
    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    estimate_chunk(0, K)
    f_ms = timed(lambda: estimate_chunk(0, K), 2) / K
    plan = torch.zeros(2, dtype=torch.float64, device="cuda")
    if rank == 0:
        fl = estimate_chunk(0, 1)[0]
        accumulate(fl)
        a_ms = timed(lambda: accumulate(fl), 4)
        plan[0], plan[1] = f_ms, a_ms
    dist.broadcast(plan, src=0)
    f_ms, a_ms = float(plan[0]), float(plan[1])
    counts = plan_round(world, Q, f_ms, a_ms)
    transport = os.environ.get("TFB200_TRANSPORT", "p2p")
    frames_per_round = sum(counts) * K
    fan = None
    if os.environ.get("TFB200_FANOUT", "1") == "1":
        from .peer import PeerFrameFanout
        fan = PeerFrameFanout(rank, world, (H, W, 3))
        fanout_box["f"] = fan
    e2e_first_round = {"j": None}

    def round_hook(j):
        # in the end-to-end pass, every non-zero rank queues the D2H of the frames it will be handed
        if io["host"] and fan is not None and rank != 0:
            if e2e_first_round["j"] is None:
                e2e_first_round["j"] = j
            lo = (j - e2e_first_round["j"]) * frames_per_round
            fan.expect(sum(1 for i in range(lo, lo + frames_per_round) if i % world == rank))
    stream = ShardedFlowStream(rank, world, K, counts, estimate_chunk, accumulate, (H, W, 2), "cuda",
                               transport=transport, round_hook=round_hook)

    def timed_rounds(first):
        stream.run(first, args.warmup)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stream.run(first + args.warmup, args.steps)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ms = timed_rounds(0)
    launches_done = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # end to end: every frame enters from pinned host memory, every RGB frame returns to it
    io["host"] = True
    ms_e2e = timed_rounds(args.warmup + args.steps)
    io["host"] = False
    launches = torch.tensor([launches_done], dtype=torch.float64, device="cuda")
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    if rank == 0:
        checksum = int(torch.from_numpy(comp.layers[0].data).long().sum())   # same for every transport
        frames = args.steps * stream.frames_per_round
        fps = frames / (float(ms) / 1000.0)
        peak, peak_src = B.measured_peak_gbs()
        n_px = H * W
        frame_bytes = fb.algorithmic_bytes(True) + (24.0 + 50.0) * n_px + 4.0 * n_px
        line = {
            "metric": "frames/sec at 4K (flow+accumulate+remap)", "value": fps, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(ms) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(B.workload_config(args), sharding=f"chunks of {K} pairs; per round {counts} chunks per rank "
                           f"(rank 0 also runs the sequential accumulate+remap); transport {transport}: "
                           + ("the producer's last post-process kernel stores the flow into rank 0's ring over NVLink "
                              "peer memory, counters + cuStreamWaitValue32 order it" if transport == "p2p" else
                              "batched NCCL send/recv, receives posted one round ahead"),
                           frames_per_step=stream.frames_per_round, state_checksum=checksum,
                           calibrated_ms={"flow_per_pair": f_ms, "accumulate_per_frame": a_ms}),
            "roofline": None,
            "e2e": {"value": frames / (float(ms_e2e) / 1000.0), "unit": "frames/s",
                    "h2d_bytes_per_step": int(stream.frames_per_round * n_px * 3 * (K + 1) / K),
                    "d2h_bytes_per_step": int(stream.frames_per_round * n_px * 3),
                    "api": "sharded stream: pinned BGR frames H2D on every rank; RGB frame i leaves through rank i % N's "
                           "PCIe link (rank 0 hands it over NVLink)" if fan is not None else
                           "sharded stream: pinned BGR frames H2D on every rank, RGB frames D2H on rank 0"},
            "gpu_launches": int(launches), "clocks": clocks,
            "pipeline_hbm_frac": frame_bytes * fps / 1e9 / (peak * world),
            "exchange_bytes_per_step": int(sum(counts[1:]) * K * n_px * 8),
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    return 0
