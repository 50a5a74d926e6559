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


# ---------------------------------------------------------------------------------------------------------------------
# sol_create_multi: ONE process, one handle over g GPUs (the form the single-threaded C++ host program uses)
# ---------------------------------------------------------------------------------------------------------------------
def _cases():
    from solaris_b200 import capi, synth
    from oraclelib import default_nebula
    return {
        "disk": lambda: (synth.massive_disk(3000, migration=True), default_nebula(), capi.RUNGE_KUTTA_FEHLBERG78),
        "trojans": lambda: (synth.trojans(5000), None, capi.DORMAND_PRINCE),
        "mixed": lambda: (synth.mixed([1, 3, 10, 200, 50, 400, 300], migration=True, seed=5), default_nebula(), capi.RUNGE_KUTTA4),
    }


@pytest.mark.parametrize("g", [2, 4, 8])
@pytest.mark.parametrize("case", ["disk", "trojans", "mixed"])
def test_multi_handle_equals_one_gpu(case, g):
    """Every entry point on a multi-GPU handle against the same call on one GPU: steps (bit-identical: a row's result
    does not depend on which GPU computes it), seam B with host pointers, downloads of sharded arrays, event counts /
    indices / records, integrals, body removal, the snapshot record."""
    if _ngpu() < g:
        pytest.skip(f"needs {g} GPUs")
    from solaris_b200 import capi
    sysm, neb, integ = _cases()[case]()
    one, many = capi.Context(0), capi.Context(n_gpus=g)
    try:
        for ctx in (one, many):
            ctx.set_frame(False); ctx.set_bodies(sysm); ctx.set_nebula(neb)
        assert many.shard_range() == (0, sysm.n)
        # seam B: Acceleration::Compute with host arrays
        a1, ag = one.compute(1.0, sysm.y0, capi.EVAL_ALL), many.compute(1.0, sysm.y0, capi.EVAL_ALL)
        assert np.array_equal(a1, ag)
        for what in (capi.RM3, capi.NN_INDEX, capi.NN_DISTANCE, capi.MIGTYPE, capi.ACCEL_GASDRAG, capi.ACCEL_MIGTYPE1, capi.ACCEL_MIGTYPE2):
            assert np.array_equal(one.download(what), many.download(what)), what
        # seam A
        for ctx in (one, many):
            ctx.upload(capi.Y0, sysm.y0)
        logs = []
        for ctx in (one, many):
            t, h, log = 0.0, 0.05, []
            for _ in range(6):
                rc, t, h, hd, att, em, ev, pr = ctx.step(integ, t, h)
                assert rc == 0, ctx.last_error()
                log.append((t, h, hd, att, em))
            logs.append(log)
        assert logs[0] == logs[1], "step-size sequence must be identical on 1 and g GPUs"
        assert np.array_equal(one.download(capi.Y0), many.download(capi.Y0))
        assert np.array_equal(one.download(capi.Y), many.download(capi.Y))
        # events: counts, merged index lists, the 120-byte records
        ev1, evg = one.detect_events(5.5, 5.2, 0.0), many.detect_events(5.5, 5.2, 0.0)
        assert len(ev1[0]) + len(ev1[1]) > 0
        for x, y in zip(ev1, evg):
            assert np.array_equal(x, y)
        assert one.event_records(12.5, 7) == many.event_records(12.5, 7)
        i1, ig = one.integrals(), many.integrals()
        scale = np.maximum(np.abs(i1), np.abs(i1[[0, 7, 7, 7, 8, 8, 8, 7, 8, 12, 12, 12, 12, 13, 14, 14]]))
        assert np.all(np.abs(ig - i1) <= 1e-12 * scale)
        # sol_run on a general system: the host loop inside the library, on every rank
        r1 = one.run(integ, t, h, 3, ejection=1.0e4)
        rg = many.run(integ, t, h, 3, ejection=1.0e4)
        assert (r1[0], r1[1].steps, r1[1].time, r1[1].h_next) == (rg[0], rg[1].steps, rg[1].time, rg[1].h_next) and r1[0] == 0
        # removal, two more steps, snapshot record
        gone = [sysm.n - 1, 300, 500] if case == "mixed" else [sysm.n - 1, 7, 40]
        for ctx in (one, many):
            ctx.remove_bodies(gone)
        t2, h2 = r1[1].time, r1[1].h_next
        for ctx in (one, many):
            t, h = t2, h2
            for _ in range(2):
                rc, t, h, *_ = ctx.step(integ, t, h)
                assert rc == 0
        assert np.array_equal(one.download(capi.Y0), many.download(capi.Y0))
        assert one.pack_phases(t) == many.pack_phases(t)
    finally:
        one.close(); many.close()


@pytest.mark.parametrize("g", [2, 4, 8])
def test_multi_handle_symmetric_kernel_against_the_exact_rows(g):
    """N = 10^5 self-gravitating bodies on g GPUs: the symmetric kernel's rounds are dealt to the ranks and the partial
    sums combined over NVLink.  1024 random sinks + the worst-conditioned ones against the extended-precision row oracle
    at the north star's 1e-13, nearest neighbours against the reference's row arithmetic; then a few RK4 steps against
    one GPU (different partial grouping -> rounding-level agreement)."""
    if _ngpu() < g:
        pytest.skip(f"needs {g} GPUs")
    from solaris_b200 import capi, synth
    from oraclelib import Oracle
    s = synth.massive_disk(100_000)
    one, many = capi.Context(0), capi.Context(n_gpus=g)
    try:
        for ctx in (one, many):
            ctx.set_frame(False); ctx.set_bodies(s); ctx.set_nebula(None)
        a = many.compute(0.0, s.y0, 0)
        nn = many.download(capi.NN_INDEX)
        o = Oracle(s, False, None)
        rng = np.random.default_rng(17)
        r2 = (s.y0[1:, :3] ** 2).sum(axis=1)
        cond = synth.GAUSS2 * (s.mass[0] + s.mass[1:]) / r2 / np.sqrt((a[1:, 3:] ** 2).sum(axis=1))
        rows = np.unique(np.concatenate([[1, s.n - 1], rng.integers(1, s.n, 1024), 1 + np.argsort(-cond)[:32]])).astype(np.int32)
        ex = o.gravity_rows_exact(s.y0, rows)
        err = np.abs(a[rows, 3:] - ex).max(axis=1) / np.sqrt((ex ** 2).sum(axis=1))
        assert err.max() <= 1.0e-13, (err.max(), int(rows[err.argmax()]))
        for i in rows[:96]:
            o.gravity_rows(s.y0, int(i), int(i) + 1, 1)
            assert nn[i] == o.side()[1][i]
        one.compute(0.0, s.y0, 0)
        assert np.array_equal(nn, one.download(capi.NN_INDEX))      # every body's neighbour equals the single-GPU result
        ys = []
        for ctx in (one, many):
            ctx.upload(capi.Y0, s.y0)
            t, h = 0.0, 0.02
            for _ in range(3):
                rc, t, h, *_ = ctx.step(capi.RUNGE_KUTTA4, t, h)
                assert rc == 0
            ys.append(ctx.download(capi.Y0))
        assert np.abs(ys[0] - ys[1]).max() <= 1e-13 * np.abs(ys[0]).max()
    finally:
        one.close(); many.close()


@pytest.mark.parametrize("g", [2, 8])
def test_dropin_program_on_several_gpus(tmp_path, g):
    """The reference's own single-threaded host program (main / Simulator / XML loader, linked unchanged) driving g GPUs
    through one sol_create_multi handle (SOLARIS_B200_GPUS=g): 20 000 self-gravitating protoplanets loaded from an XML
    file, RK4, against the same program on one GPU.  Same snapshots (the symmetric kernel groups its partial sums
    differently on g GPUs: 1e-12), and the state really lives on g devices (stats line)."""
    if _ngpu() < g:
        pytest.skip(f"needs {g} GPUs")
    import re
    import struct
    import xmlgen
    from solaris_b200 import synth
    dropin = os.path.join(ROOT, "solaris_b200", "host", "_build", "solaris_b200_dropin")
    if not os.path.exists(dropin):
        pytest.skip("prebuilt drop-in program missing")
    s = synth.massive_disk(20_000)
    bodies = []
    for k in range(1, s.n):
        y = [float(v) for v in s.y0[k]]
        bodies.append(f'        <Body type="protoplanet" name="p{k}">\n          <Phase>\n'
                      f'            <Position x="{y[0]!r}" y="{y[1]!r}" z="{y[2]!r}" unit="au" />\n'
                      f'            <Velocity x="{y[3]!r}" y="{y[4]!r}" z="{y[5]!r}" unit="auday" />\n          </Phase>\n'
                      f'          <Characteristics>\n            <Mass value="{float(s.mass[k])!r}" unit="solar" />\n          </Characteristics>\n        </Body>\n')
    xml = xmlgen.make("disk on several GPUs", "RungeKutta4", "0.002", "0.001", bodies)
    outs = {}
    for n in (1, g):
        d = tmp_path / f"g{n}"
        d.mkdir()
        (d / "in.xml").write_text(xml)
        env = dict(os.environ, OSTYPE="linux", SOLARIS_B200_GPUS=str(n), SOLARIS_B200_STATS="1")
        r = subprocess.run([dropin, "-i", str(d / "in.xml")], cwd=str(d), env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        b = (d / "Phases.dat").read_bytes()
        snaps, off = [], 0
        while off < len(b):
            t, nb = struct.unpack_from("<di", b, off); off += 12
            rec = np.frombuffer(b, dtype=np.dtype([("id", "<i4"), ("y", "<f8", (6,))]), count=nb, offset=off); off += 52 * nb
            snaps.append((t, rec["id"].copy(), rec["y"].copy()))
        outs[n] = (snaps, r.stderr)
    one, many = outs[1][0], outs[g][0]
    assert len(one) == len(many) >= 3 and len(one[0][1]) == s.n
    for (t1, id1, y1), (tg, idg, yg) in zip(one, many):
        assert t1 == tg and np.array_equal(id1, idg)
        assert np.abs(y1 - yg).max() <= 1e-12 * np.abs(y1).max()
    assert np.abs(one[-1][2] - one[0][2]).max() > 0
    assert re.search(r"(\d+) steps", outs[g][1])
