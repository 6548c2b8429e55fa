"""CPU test of the N > 1 path: two gloo ranks each run the tick on their contiguous shard of the SAME global synthetic
batch; the all-gathered solutions equal the single-process result (rank-count invariance).  The compute
on each rank is the oracle here (no GPU in this container); on the B200 box bench.py runs the same host logic with
the CUDA path and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H

NAME, N_GLOBAL, SEED, SIGMA = "a1_px100_pin_ver", 7, 20260001, 5e-3      # 7 states over 2 ranks: ragged shards


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _solve(q, targets, mem, ref):
    rm = H.make_oracle(NAME)
    rm.setTasks(Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=True, Joint=True)
    rm.setConstraints(CoM=False, Trunk=True, FR=True, FL=True, RR=True, RL=True, Grip=False)
    x, st = [], []
    for s in range(q.shape[0]):
        r = H.oracle_step_one(rm, q[s], targets[s], mem[s], ref[s], tail=False)
        x.append(r["qdot"])
        st.append(r["status"])
    return np.array(x).reshape(q.shape[0], -1), np.array(st, dtype=np.int32)


def _worker(rank, world, port, inputs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wbc_b200 import sharding
    q, targets, mem, ref = inputs
    lo, hi = sharding.shard_range(N_GLOBAL, rank, world)
    x, st = _solve(q[lo:hi], targets[lo:hi], mem[lo:hi], ref[lo:hi])
    gx = sharding.gather_states(torch.from_numpy(x), N_GLOBAL)
    gs = sharding.gather_states(torch.from_numpy(st), N_GLOBAL)
    t = sharding.max_over_ranks(1.0 + rank, torch.device("cpu"))
    if rank == 0:
        np.save(os.path.join(out_dir, "x.npy"), gx.numpy())
        np.save(os.path.join(out_dir, "st.npy"), gs.numpy())
        np.save(os.path.join(out_dir, "t.npy"), np.array([t]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    from wbc_b200 import sharding
    for n in (0, 1, 7, 4096, 65536, 1 << 20):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = sharding.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    assert sharding.shard_range(1 << 20, 3, 8) == (3 * 131072, 4 * 131072)       # BASELINE config 4 at 8 GPUs
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_two_rank_gloo_run_equals_single_process(tmp_path):
    import bench
    inputs = bench.cpu_inputs(NAME, N_GLOBAL, SEED, SIGMA)
    x_ref, st_ref = _solve(*inputs)
    mp.spawn(_worker, args=(2, _free_port(), inputs, str(tmp_path)), nprocs=2, join=True)
    gx, gs = np.load(tmp_path / "x.npy"), np.load(tmp_path / "st.npy")
    assert gx.shape == x_ref.shape and np.array_equal(gx, x_ref)       # same arithmetic on the same inputs: bit-equal
    assert np.array_equal(gs, st_ref) and (gs == 0).all()
    assert float(np.load(tmp_path / "t.npy")[0]) == 2.0                # timing reduction is a max over ranks
