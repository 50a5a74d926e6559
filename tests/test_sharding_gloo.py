"""N>1 path on CPU: world_size-2 gloo processes reproduce the multi-GPU plumbing of the device library
(csrc/api.cu: eval_force / exchange_sources / read_error_max) over the ORACLE's row-subset kernels:
sinks are sharded with the library's own partition rule (sol_shard_of), every RK4 stage all-gathers the
trial state of the shards, and a max-norm is all-reduced.  The sharded result must equal the unsharded
oracle step bit for bit (rows are independent, max is associative)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bary, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from solaris_b200 import capi, synth
    from oraclelib import Oracle

    s = synth.mixed([1, 3, 0, 40, 0, 0, 57], migration=False, seed=21)
    if bary:
        s = synth.to_barycentric(s)
    n = s.n
    lo, hi = capi.shard_of(n, world, rank)
    o = Oracle(s, bary, None)

    def gather_rows(local_rows):
        """all-gather of ragged shards == the grouped ncclBroadcast in exchange_sources()"""
        full = np.zeros((n, 6))
        for r in range(world):
            rlo, rhi = capi.shard_of(n, world, r)
            buf = torch.from_numpy(local_rows.copy() if r == rank else np.zeros((rhi - rlo, 6)))
            dist.broadcast(buf, src=r)
            full[rlo:rhi] = buf.numpy()
        return full

    def f_rows(y):
        return o.gravity_rows(y, lo, hi, 1)          # this rank's sinks against ALL sources

    h = 0.37
    y0 = s.y0.copy()
    k1 = f_rows(y0)
    y = gather_rows(y0[lo:hi] + h * (0.5 * k1))
    k2 = f_rows(y)
    y = gather_rows(y0[lo:hi] + h * (0.5 * k2))
    k3 = f_rows(y)
    y = gather_rows(y0[lo:hi] + h * (1.0 * k3))
    k4 = f_rows(y)
    ynew = gather_rows(y0[lo:hi] + h * (1.0 / 6.0 * k1 + 1.0 / 3.0 * k2 + 1.0 / 3.0 * k3 + 1.0 / 6.0 * k4))
    # max-norm all-reduce (read_error_max): the bit pattern of a non-negative double orders like an int
    local = np.abs(k1 - k4).max() if hi > lo else 0.0
    t = torch.tensor([np.float64(local).view(np.int64)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    emax = np.int64(t.item()).view(np.float64)
    if rank == 0:
        np.savez(os.path.join(out_dir, f"sharded_{int(bary)}.npz"), y=ynew, emax=emax, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("bary", [False, True])
def test_two_rank_sharded_rk4_equals_unsharded(tmp_path, bary):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), bary, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / f"sharded_{int(bary)}.npz")
    sys.path.insert(0, HERE)
    from solaris_b200 import synth
    from oraclelib import Oracle
    s = synth.mixed([1, 3, 0, 40, 0, 0, 57], migration=False, seed=21)
    if bary:
        s = synth.to_barycentric(s)
    o = Oracle(s, bary, None)
    k1 = o.compute(0.0, s.y0, 0)
    r, t, hn, hd, _, _ = o.step(1, 0.0, 0.37)
    assert r == 0
    assert np.array_equal(got["y"], o.array("y0")), "sharded RK4 step must equal the unsharded one bit for bit"
    assert 0 < int(got["lo"]) or int(got["hi"]) < s.n
    # emax: recompute unsharded
    o2 = Oracle(s, bary, None)
    k1 = o2.compute(0.0, s.y0, 0)
    y = s.y0 + 0.37 * (0.5 * k1); k2 = o2.compute(0.0, y, 0)
    y = s.y0 + 0.37 * (0.5 * k2); k3 = o2.compute(0.0, y, 0)
    y = s.y0 + 0.37 * (1.0 * k3); k4 = o2.compute(0.0, y, 0)
    assert float(got["emax"]) == np.abs(k1 - k4).max()
