"""CPU (gloo, world_size > 1): the host logic of the frame-pair sharding -- round plan, chunk
ownership, and that rank 0 accumulates every frame exactly once, in frame order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from transflow_b200.distributed import ShardedFlowStream, plan_rank0_pairs, plan_round, round_slots


def test_plan_round_balances_rank0():
    assert plan_round(1, 4, 1.0, 0.2) == [4]
    # a = 0.15: c0 = 4 * (1 - (N-1) * 0.15) / 1.15
    assert plan_round(2, 4, 1.0, 0.15) == [2, 4]
    assert plan_round(4, 4, 1.0, 0.15) == [1, 4, 4, 4]
    assert plan_round(8, 4, 1.0, 0.15) == [0] + [4] * 7
    assert plan_round(2, 4, 1.0, 0.0) == [4, 4]
    for world in (2, 3, 8):
        for a in (0.0, 0.1, 0.5, 2.0):
            c = plan_round(world, 4, 1.0, a)
            total = sum(c)
            # rank 0 never finishes later than a producer
            assert c[0] * 1.0 + total * a <= 4 * 1.0 + 1e-9 or c[0] == 0
    assert round_slots([2, 4, 4]) == [0, 0, 1, 1, 1, 1, 2, 2, 2, 2]


def test_plan_rank0_pairs():
    # p0 = q k (F - (N - 1) A) / (F + A), floored and clamped
    assert plan_rank0_pairs(1, 4, 8, 1.0, 0.1) == 32
    assert plan_rank0_pairs(2, 4, 8, 0.925, 0.1026) == 25
    assert plan_rank0_pairs(4, 4, 8, 0.925, 0.1026) == 19
    assert plan_rank0_pairs(8, 4, 8, 0.925, 0.1026) == 6
    assert plan_rank0_pairs(8, 4, 8, 0.5, 0.1) == 0            # rank 0 saturated by the accumulation alone
    for world in (2, 4, 8):
        for a in (0.0, 0.05, 0.2, 1.0):
            p0 = plan_rank0_pairs(world, 4, 8, 1.0, a)
            assert 0 <= p0 <= 32
            assert p0 * 1.0 + (p0 + (world - 1) * 32) * a <= 32 * 1.0 + 1e-9 or p0 == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, counts, k, rounds, out_dir, rank0_pairs=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def estimate_chunk(first_pair, n_pairs, join=True):
        # the "flow" of pair t is a tiny tensor filled with t (and tagged with the producing rank)
        return [torch.tensor([[[float(first_pair + i), float(rank)]]]) for i in range(n_pairs)]

    def accumulate(flow):
        seen.append((int(flow[0, 0, 0]), int(flow[0, 0, 1])))

    stream = ShardedFlowStream(rank, world, k, counts, estimate_chunk, accumulate, (1, 1, 2), "cpu",
                               rank0_pairs=rank0_pairs)
    for r in range(rounds):
        stream.run_round(r)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "seen.npy"), np.asarray(seen))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,counts", [(2, [2, 4]), (2, [0, 3]), (3, [1, 2, 2])])
def test_sharded_stream_orders_frames(world, counts, tmp_path):
    k, rounds = 3, 3
    mp.spawn(_worker, args=(world, _free_port(), counts, k, rounds, str(tmp_path)), nprocs=world, join=True)
    seen = np.load(tmp_path / "seen.npy")
    total = rounds * sum(counts) * k
    assert seen.shape == (total, 2)
    np.testing.assert_array_equal(seen[:, 0], np.arange(total))      # frame order, each exactly once
    owners = round_slots(counts)
    expect_rank = [owners[(t // k) % len(owners)] for t in range(total)]
    np.testing.assert_array_equal(seen[:, 1], expect_rank)           # produced by the planned owner


@pytest.mark.parametrize("world,counts,p0", [(2, [0, 3], 5), (3, [0, 2, 2], 1), (2, [0, 2], 0)])
def test_sharded_stream_with_rank0_pairs_at_the_end_of_the_round(world, counts, p0, tmp_path):
    """rank0_pairs mode: the producers' chunks come first, rank 0's own chunk of p0 pairs closes the round."""
    k, rounds = 3, 3
    mp.spawn(_worker, args=(world, _free_port(), counts, k, rounds, str(tmp_path), p0), nprocs=world, join=True)
    seen = np.load(tmp_path / "seen.npy")
    per_round = sum(counts[1:]) * k + p0
    total = rounds * per_round
    assert seen.shape == (total, 2)
    np.testing.assert_array_equal(seen[:, 0], np.arange(total))
    owners = []
    for r, c in enumerate(counts):
        if r > 0:
            owners += [r] * (c * k)
    owners += [0] * p0
    np.testing.assert_array_equal(seen[:, 1], [owners[t % per_round] for t in range(total)])
