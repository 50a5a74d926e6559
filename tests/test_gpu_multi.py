"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): two ranks, sinks sharded, sources
all-gathered over NCCL per evaluation.  The sharded trajectory must equal the single-GPU one bit for
bit (row results do not depend on which GPU computes them; the indirect sum is recomputed identically)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["SOL_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SOL_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
from solaris_b200 import capi, synth
from oraclelib import default_nebula
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)
case = os.environ["SOL_CASE"]
if case == "disk":
    s, neb, integ = synth.massive_disk(3000, migration=True), default_nebula(), capi.RUNGE_KUTTA_FEHLBERG78
elif case == "bigdisk":
    s, neb, integ = synth.massive_disk(9000), None, capi.RUNGE_KUTTA4
elif case == "trojans":
    s, neb, integ = synth.trojans(5000), None, capi.DORMAND_PRINCE
else:
    s, neb, integ = synth.mixed([1, 3, 10, 200, 50, 400, 300], migration=True, seed=5), default_nebula(), capi.RUNGE_KUTTA4
ctx = capi.Context(rank)
uid = [capi.Context.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.dist_init(rank, world, uid[0])
ctx.set_pair_algorithm(1)        # symmetric kernel from 4096 bodies ("bigdisk" exercises its multi-GPU path)
ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(neb)
t, h = 0.0, 0.05
log = []
for _ in range(6):
    rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
    assert rc == 0, ctx.last_error()
    log.append((t, h, hd, att, em))
ctx.gather_state()
y = ctx.download(capi.Y0)
ej, hc, co = ctx.detect_events(5.5, 5.2, 0.0)
cnt = np.zeros(3, dtype=np.int32)
ctx._check(ctx.lib.sol_detect_events(ctx.h, 5.5, 5.2, 0.0, cnt.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_int))))
integ16 = ctx.integrals()
lo_hi = np.array(ctx.shard_range())
nn, nnd = ctx.download(capi.NN_INDEX), ctx.download(capi.NN_DISTANCE)
# removal on every rank (gathers, compacts, re-shards), then the shrunk system steps on; snapshot record of the result
y2 = rec = None
if case != "bigdisk":
    # (mixed: planetesimals and a test particle - removing a body AHEAD of the drag class would hand a planetesimal
    #  the cD = 0 slot of a massive body, the reference's own quirk, and the NaN drag that follows)
    ctx.remove_bodies([s.n - 1, 300, 500] if case == "mixed" else [s.n - 1, 7, 40])
    for _ in range(2):
        rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
        assert rc == 0, ctx.last_error()
    rec = np.frombuffer(ctx.pack_phases(t), dtype=np.uint8)       # gathers the state first
    y2 = ctx.download(capi.Y0)
if rank == 0:
    extra = {} if y2 is None else {"y2": y2, "rec": rec, "t2": np.float64(t)}
    np.savez(os.environ["SOL_OUT"], y=y, log=np.array(log), lo_hi=lo_hi, nn=nn, nnd=nnd, ev_counts=cnt, ej_local=ej,
             integrals=integ16, **extra)
dist.barrier(); dist.destroy_process_group()
'''


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("case", ["disk", "trojans", "mixed"])
def test_two_gpus_equal_one_gpu(tmp_path, case):
    from solaris_b200 import capi, synth
    from oraclelib import default_nebula
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "multi.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SOL_ROOT=ROOT, SOL_CASE=case, SOL_OUT=out)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    got = np.load(out)
    if case == "disk":
        sysm, neb, integ = synth.massive_disk(3000, migration=True), default_nebula(), capi.RUNGE_KUTTA_FEHLBERG78
    elif case == "trojans":
        sysm, neb, integ = synth.trojans(5000), None, capi.DORMAND_PRINCE
    else:
        sysm, neb, integ = synth.mixed([1, 3, 10, 200, 50, 400, 300], migration=True, seed=5), default_nebula(), capi.RUNGE_KUTTA4
    ctx = capi.Context(0)
    ctx.set_frame(False); ctx.set_bodies(sysm); ctx.set_nebula(neb)
    t, h = 0.0, 0.05
    log = []
    for _ in range(6):
        rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
        assert rc == 0
        log.append((t, h, hd, att, em))
    assert np.array_equal(np.array(log), got["log"]), "step-size sequence must be identical on 1 and 2 GPUs"
    assert np.array_equal(ctx.download(capi.Y0), got["y"]), "sharded state must equal the single-GPU state bit for bit"
    assert 0 < got["lo_hi"][1] < sysm.n
    # event detection: global counts equal the single-GPU counts, rank 0 holds the candidates of its own shard
    ej1, hc1, co1 = ctx.detect_events(5.5, 5.2, 0.0)
    assert list(got["ev_counts"]) == [len(ej1), len(hc1), len(co1)]
    lo, hi = got["lo_hi"]
    assert np.array_equal(got["ej_local"], ej1[(ej1 >= lo) & (ej1 < hi)])
    # device integrals (potential energy over both shards, all-reduced)
    i1 = ctx.integrals()
    scale = np.maximum(np.abs(i1), np.abs(i1[[0, 7, 7, 7, 8, 8, 8, 7, 8, 12, 12, 12, 12, 13, 14, 14]]))
    assert np.all(np.abs(got["integrals"] - i1) <= 1e-12 * scale)
    # body removal on the sharded context (gather, compaction, re-sharding), two more steps, snapshot record
    ctx.remove_bodies([sysm.n - 1, 300, 500] if case == "mixed" else [sysm.n - 1, 7, 40])
    for _ in range(2):
        rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
        assert rc == 0
    assert t == float(got["t2"])
    y2 = ctx.download(capi.Y0)
    assert np.isfinite(y2).all()
    assert np.array_equal(y2, got["y2"])
    assert ctx.pack_phases(t) == got["rec"].tobytes()
    ctx.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpus_symmetric_kernel(tmp_path):
    """9000 self-gravitating bodies: the symmetric kernel's rounds are dealt to the two ranks and the
    partial sums all-reduced.  Summation order differs from one GPU -> agreement to rounding, not bits."""
    from solaris_b200 import capi, synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "multi.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SOL_ROOT=ROOT, SOL_CASE="bigdisk", SOL_OUT=out)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    got = np.load(out)
    sysm = synth.massive_disk(9000)
    ctx = capi.Context(0)
    ctx.set_pair_algorithm(1)
    ctx.set_frame(False); ctx.set_bodies(sysm); ctx.set_nebula(None)
    t, h = 0.0, 0.05
    for _ in range(6):
        rc, t, h, hd, att, em, ev, pr = ctx.step(capi.RUNGE_KUTTA4, t, h)
        assert rc == 0
    y1 = ctx.download(capi.Y0)
    # rank 0 only has the NN arrays of its own shard
    lo, hi = got["lo_hi"]
    assert np.array_equal(ctx.download(capi.NN_INDEX)[lo:hi], got["nn"][lo:hi])
    assert np.abs(y1 - got["y"]).max() <= 1e-13 * np.abs(y1).max()
    ctx.close()
