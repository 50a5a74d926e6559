"""The C-ABI shared library: loads without a GPU, exports every symbol include/solaris_b200.h declares,
and fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from solaris_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "solaris_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sol_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = header_functions()
    assert len(names) >= 25
    for nm in names:
        assert hasattr(lib, nm), f"{nm} is declared in include/solaris_b200.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS out of sync with the header"


def test_product_does_not_link_the_oracle():
    """Nothing under solaris_b200/ may reference oracle/ (the oracle is a checker, never the product)."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "solaris_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"liboracle|oraclelib|libref_harness|oracle_compute|oracle_step", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_shard_partition_rule():
    for n, g in ((1_000_003, 8), (262_144, 4), (9, 2), (100, 8), (31, 3)):
        prev = 0
        for r in range(g):
            lo, hi = capi.shard_of(n, g, r)
            assert lo == prev and lo <= hi <= n
            assert (hi - lo) % 32 == 0 or hi == n
            prev = hi
        assert prev == n
    assert capi.shard_of(10, 1, 0) == (0, 10)


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="needs a machine WITHOUT a CUDA device")
def test_fails_loudly_without_gpu():
    with pytest.raises(capi.SolarisError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.sol_create(0, C.byref(h)) == 1 and not h.value


def test_symmetric_kernel_schedule_covers_every_block_pair_once():
    """Host logic of the symmetric pair kernel: rounds x CTAs -> block pairs; each unordered pair of
    blocks exactly once, each diagonal block once; the per-rank round ranges tile the rounds."""
    lib = capi.load_library()
    for nb in (1, 2, 3, 4, 7, 8, 16, 17, 33, 64):
        seen = {}
        for r in range(nb // 2 + 1):
            for p in range(nb):
                q = C.c_int(-1)
                ok = lib.sol_sym_round_pair(nb, r, p, C.byref(q))
                assert ok in (0, 1)
                if ok:
                    key = (min(p, q.value), max(p, q.value))
                    seen[key] = seen.get(key, 0) + 1
                    assert (r == 0) == (p == q.value)
        want = {(a, b) for a in range(nb) for b in range(a, nb)}
        assert set(seen) == want and all(v == 1 for v in seen.values()), nb
        for g in (1, 2, 3, 8):
            prev = 0
            for rank in range(g):
                lo, hi = C.c_int(0), C.c_int(0)
                assert lib.sol_sym_rounds_of_rank(nb, g, rank, C.byref(lo), C.byref(hi)) == 0
                assert lo.value == prev and hi.value >= lo.value
                prev = hi.value
            assert prev == nb // 2 + 1
        # the cost-balanced split the library uses: every (round, CTA) with work belongs to exactly one rank, the shares
        # are contiguous in the (round, CTA) sequence and differ by at most two CTAs in cost
        for g in (1, 2, 3, 5, 8):
            owner = {}
            costs = []
            for rank in range(g):
                out = (C.c_int * 4)()
                assert lib.sol_sym_work_of_rank(nb, g, rank, out) == 0
                r0, p0, r1, p1 = list(out)
                cost = 0.0
                for r in range(r0, r1 + 1):
                    for p in range(nb):
                        if (r == r0 and p < p0) or (r == r1 and p >= p1):
                            continue
                        q = C.c_int(-1)
                        if lib.sol_sym_round_pair(nb, r, p, C.byref(q)) == 1:
                            assert (r, p) not in owner
                            owner[(r, p)] = rank
                            cost += 0.75 if r == 0 else 1.0
                costs.append(cost)
            assert len(owner) == nb * (nb + 1) // 2, (nb, g)
            seq = [owner[k] for k in sorted(owner)]
            assert seq == sorted(seq)
            assert max(costs) - min(costs) <= 2.0, (nb, g, costs)      # within two CTAs of each other (978 rounds x 1954 CTAs at N = 10^6)
    q = C.c_int(0)
    assert lib.sol_sym_round_pair(4, 3, 0, C.byref(q)) == -1


def test_pair_launch_plan():
    """Launch plan of the ordered pair kernel (pure host logic): the chunks cover the sources exactly once with at most 32
    partial sums per sink; at most 256 sources are never cut (the single-CTA kernel's summation order, asserted bit-identical
    on the GPU); a mid-size launch is one wave of at most 288 CTAs and is cut from the GLOBAL sink count alone, so that a
    sharded context sums in the same order as an unsharded one."""
    lib = capi.load_library()

    def plan(ni, nj, ni_all=0):
        out = (C.c_int * 3)()
        assert lib.sol_plan_pairs(ni, nj, ni_all, out) == 0
        return tuple(out)

    for ni in (1, 100, 257, 1000, 2999, 12000, 100000, 400000, 1000000):
        for nj in (1, 31, 256, 257, 300, 1000, 2999, 8000, 20000, 1000000):
            I, splits, chunk = plan(ni, nj)
            assert I in (1, 2, 4) and 1 <= splits <= 32 and chunk >= 1
            assert splits * chunk >= nj and (splits - 1) * chunk < nj, (ni, nj, splits, chunk)
            if nj <= 256:
                assert splits == 1
    for n in (300, 1000, 2000, 3000, 4000):            # mid-size: every CTA in one wave, chunks of at least 32 sources
        I, splits, chunk = plan(n, n)
        iblocks = -(-n // 128)
        assert I == 1 and iblocks * splits <= 288 and chunk >= 32, (n, splits, chunk)
        for ranks in (2, 3, 8):
            assert plan(-(-n // ranks), n, n)[1:] == (splits, chunk), (n, ranks)
    assert lib.sol_plan_pairs(0, 10, 0, (C.c_int * 3)()) != 0


def test_missing_library_is_a_loud_error(monkeypatch, tmp_path):
    """No silent fallback when the CUDA library has not been built."""
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", str(tmp_path / "libsolaris_b200.so"))
    with pytest.raises(RuntimeError) as e:
        capi.load_library()
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(RuntimeError):
        capi.Context(0)


def test_sass_uses_bulk_copy_mbarrier_and_rsq64h():
    """The shipped cubin really contains the sm_100a instructions DESIGN.md claims: UBLKCP (cp.async.bulk),
    SYNCS (mbarrier transaction counting) and MUFU.RSQ64H (fp64 rsqrt seed)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UBLKCP", "SYNCS.ARRIVE.TRANS64", "MUFU.RSQ64H", "SHFL.IDX", "DFMA"):
        assert mnemonic in sass, mnemonic


def test_generated_header_is_up_to_date():
    """csrc/ilp_asm.cuh is generated (tools/gen_ilp_asm.py): the committed file is what the generator writes."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_ilp_asm", os.path.join(root, "tools", "gen_ilp_asm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert open(mod.PATH).read() == mod.generate()
