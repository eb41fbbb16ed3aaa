"""CPU tests of the host logic: C-ABI surface, data model, the generated recurrences (through
the test-only host build), the Boys restatement, error behaviour without a GPU, and the
world_size-2 gloo path of the J/K all-reduce."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden


def test_abi_exports_every_declared_symbol(rc):
    hdr = open(os.path.join(ROOT, "include", "rchem_eri.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(rchem_[a-z_0-9A-Z]+|coulomb_repulsion)\s*\(", hdr))
    assert {"rchem_build_I", "rchem_jk_direct", "rchem_jk_inmem", "coulomb_repulsion",
            "rchem_basis_new", "rchem_schwarz", "rchem_quartet_list"} <= names
    out = subprocess.run(["nm", "-D", "--defined-only", rc.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = names - exported
    assert not missing, f"declared in include/rchem_eri.h but not exported: {missing}"
    assert rc._lib.rchem_version() >= 100


def test_library_has_sm100a_code(rc):
    out = subprocess.run(["cuobjdump", "-lelf", rc.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


# (shell counts: the sp shells of STO-3G / 6-31G stay fused -- 4 and 7 shells, as in the Basis Set
# Exchange data; 6-31G* has a d shell and keeps s and p segmented)
@pytest.mark.parametrize("name,nbf,nshell", [("STO-3G", 7, 4), ("6-31G", 13, 7), ("6-31G*", 19, 10)])
def test_basis_new_matches_oracle_statement(rc, orc, geo, name, nbf, nshell):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = rc.Basis.new(z, x, name)
    ob = orc.make_basis(z, x, name)
    assert len(b) == nbf == ob.n
    origins, powers, off, exps, coefs, norms = b.export()
    assert np.array_equal(origins, ob.origins) and np.array_equal(powers, ob.powers)
    assert np.array_equal(off, ob.prim_offset) and np.array_equal(exps, ob.exps)
    assert np.array_equal(coefs, ob.coefs)
    assert np.abs(norms / ob.norms - 1).max() < 4e-16
    l, first = b.shells()
    assert len(l) == nshell
    assert sum(rc.shell_nfun(v) for v in l) == nbf
    b.set_fuse_sp(False)  # RCHEM_OPT_FUSE_SP = 0: s and p parts as separate shells
    assert len(b.shells()[0]) == {"STO-3G": 5, "6-31G": 9, "6-31G*": 10}[name]
    # function order atom -> shell -> am -> component (basis.rs:186-203): O first, then H, H
    assert np.allclose(origins[0], x[0]) and np.allclose(origins[-1], x[2])


def test_basis_from_cgtos_roundtrip_and_layout_errors(rc, orc, geo):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    ob = orc.make_basis(z, x, "6-31G*")
    b = rc.Basis.from_cgtos(ob.origins, ob.powers, ob.prim_offset, ob.exps, ob.coefs, ob.norms)
    assert len(b) == 19
    # drop one p component: no longer complete Cartesian shells
    keep = [i for i in range(ob.n) if i != 3]
    off = [0]
    ex, co, no = [], [], []
    for i in keep:
        s, e = ob.prim_offset[i], ob.prim_offset[i + 1]
        ex += list(ob.exps[s:e]); co += list(ob.coefs[s:e]); no += list(ob.norms[s:e])
        off.append(len(ex))
    with pytest.raises(rc.RchemError) as ei:
        rc.Basis.from_cgtos(ob.origins[keep], ob.powers[keep], off, ex, co, no)
    assert ei.value.code == -3
    with pytest.raises(rc.RchemError) as ei:
        rc.Basis.new(z, x, "cc-pVQZ")
    assert ei.value.code == -8
    with pytest.raises(rc.RchemError):
        rc.Basis.new(np.array([8, 1, 79], dtype=np.uint64), x, "STO-3G")


def test_helpers_match_reference_semantics(rc, orc):
    assert rc.get_ijk_list(3).tolist() == orc.ijk_list(3).tolist()  # shell.rs:44-61
    for pw, a in (([0, 0, 0], 1.3), ([1, 0, 0], 0.4), ([1, 1, 0], 0.8), ([2, 0, 0], 0.8), ([0, 1, 2], 5.0)):
        ref = orc.lib().orc_normalization(np.array(pw, dtype=np.int32), a)
        assert abs(rc.normalization(pw, a) / ref - 1) < 4e-16
    g = golden("ijkl_ref.npz")
    assert [rc.ijkl2intindex(*map(int, q)) for q in g["ijkl"]] == g["index"].tolist()


def test_no_cpu_fallback(rc, geo):
    if rc.device_count() > 0:
        pytest.skip("a CUDA device is present")
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = rc.Basis.new(z, x, "STO-3G")
    J, K = np.zeros((7, 7)), np.zeros((7, 7))
    with pytest.raises(rc.RchemError) as ei:
        rc.JK_direct(J, K, b, np.eye(7))
    assert ei.value.code == -5
    with pytest.raises(rc.RchemError):
        rc.build_I(b)
    with pytest.raises(rc.RchemError):
        rc.JK_inmem(np.zeros((2, 2, 2, 2)), np.eye(2))
    with pytest.raises(rc.RchemError):
        rc.coulomb_repulsion(1.1, 1.2, 1.3, 1.4, [1, 0, 1], [0, 1, 2], [0, 0, 3], [0, 0, 4], 1, 1, 1, 1, [0] * 12)


def test_argument_validation(rc, geo):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = rc.Basis.new(z, x, "STO-3G")
    A = np.arange(49, dtype=np.float64).reshape(7, 7)
    J, K = np.zeros((7, 7)), np.zeros((7, 7))
    if rc.device_count() == 0:  # any D is accepted now; without a device the call still fails loudly
        with pytest.raises(rc.RchemError) as ei:
            rc.JK_direct(J, K, b, A)
        assert ei.value.code == -5
    for key, val in ((rc.OPT_HEAVY_PASSES, 0.0), (rc.OPT_HEAVY_PASSES, 1e9), (rc.OPT_LIGHT_KERNEL, 0.0),
                     (rc.OPT_SYMMETRIC_D_ONLY, 1.0)):
        rc._check(rc._lib.rchem_set_option(b._h, key, val))
        assert rc._lib.rchem_get_option(b._h, key) == val
    with pytest.raises(rc.RchemError):
        b.set_heavy_passes(-1.0)
    with pytest.raises(ValueError):
        rc.JK_direct(np.zeros((6, 6)), K, b, np.eye(7))
    with pytest.raises(rc.RchemError):
        b.set_boys(3)
    with pytest.raises(rc.RchemError):
        b.set_schwarz_tau(-1.0)
    assert rc._lib.rchem_get_option(b._h, rc.OPT_FAR_SCHED) == 1.0  # far-field scheduling: on
    b.set_far_sched(False)
    assert rc._lib.rchem_get_option(b._h, rc.OPT_FAR_SCHED) == 0.0
    with pytest.raises(rc.RchemError):
        rc._check(rc._lib.rchem_set_option(b._h, rc.OPT_FAR_SCHED, 2.0))
    with pytest.raises(rc.RchemError) as ei:  # g function: tier-1 kernel stops at f
        rc.coulomb_repulsion_batch(np.zeros((1, 12)), np.ones((1, 4)),
                                   np.array([[4, 0, 0] + [0] * 9]), np.ones((1, 4)))
    assert ei.value.code == -2


# ---- the generated recurrences, on the host -----------------------------------------------
def _shell_blocks(hostcheck, ob, boys, I_ref):
    ls = np.zeros(ob.n, dtype=np.int32)
    bf = np.zeros(ob.n, dtype=np.int32)
    ns = hostcheck.hostcheck_nshells(*ob.args(), ls, bf)
    nc = lambda l: (l + 1) * (l + 2) // 2
    worst, classes = 0.0, set()
    out = np.zeros(1296)
    for sa in range(ns):
        for sb in range(ns):
            if ls[sa] < ls[sb]:
                continue
            for sc in range(ns):
                for sd in range(ns):
                    if ls[sc] < ls[sd] or (ls[sa], ls[sb]) < (ls[sc], ls[sd]):
                        continue
                    n = hostcheck.hostcheck_shell_quartet(*ob.args(), sa, sb, sc, sd, boys, out)
                    assert n > 0
                    blk = I_ref[bf[sa]:bf[sa] + nc(ls[sa]), bf[sb]:bf[sb] + nc(ls[sb]),
                                bf[sc]:bf[sc] + nc(ls[sc]), bf[sd]:bf[sd] + nc(ls[sd])]
                    worst = max(worst, np.abs(out[:n].reshape(blk.shape) - blk).max())
                    classes.add((ls[sa], ls[sb], ls[sc], ls[sd]))
    return worst, classes


def test_all_21_classes_against_oracle(hostcheck, orc, geo, ref_or_restated):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    ob = orc.make_basis(z, x, "6-31G*")
    with ref_or_restated():
        I_ref = orc.build_I(ob)
    worst, classes = _shell_blocks(hostcheck, ob, 0, I_ref)
    assert len(classes) == 21
    assert worst < 1e-12, worst  # reference Boys: the north-star tolerance
    I_x = orc.build_I(ob, orc.BOYS_EXACT)
    worst, _ = _shell_blocks(hostcheck, ob, 1, I_x)
    assert worst < 1e-13, worst  # exact Boys vs the exact-Boys twin
    assert 1e-9 < np.abs(I_x - I_ref).max() < 2e-8  # the two flavours differ by the Boys error


def test_far_field_form_all_21_classes(hostcheck, orc, geo):
    """primitive_quartet_far (eri_core.h): for shell quartets whose every primitive has
    x >= 48 the rho-free point-multipole form must equal the general code (exact Boys), which
    takes its asymptotic branch there; and the bounding-sphere proof the block kernel uses
    must never call a quartet far that has a primitive with x < 48."""
    z1, x1 = geo.molecule(geo.WATER_CRAWFORD)
    rng = np.random.default_rng(7)
    seen = set()
    worst = 0.0
    n_proved = n_far = n_total = 0
    for shift in ([9.0, 0.0, 0.0], [14.0, 3.0, -2.0], [25.0, -11.0, 7.0], [60.0, 40.0, 10.0]):
        z = np.concatenate([z1, z1])
        x = np.concatenate([x1, x1 @ np.linalg.qr(rng.standard_normal((3, 3)))[0] + np.array(shift)])
        ob = orc.make_basis(z, x, "6-31G*")
        ls = np.zeros(ob.n, dtype=np.int32)
        bf = np.zeros(ob.n, dtype=np.int32)
        ns = hostcheck.hostcheck_nshells(*ob.args(), ls, bf)
        half = ns // 2
        far_out, gen_out = np.zeros(1296), np.zeros(1296)
        quartets = [(sa, sb, sc, sd) for sa in range(half) for sb in range(half)
                    for sc in range(half, ns) for sd in range(half, ns)
                    if ls[sa] >= ls[sb] and ls[sc] >= ls[sd] and (ls[sa], ls[sb]) >= (ls[sc], ls[sd])]
        per_class = {}
        for i in rng.permutation(len(quartets)):  # up to 60 quartets of every class
            sa, sb, sc, sd = quartets[i]
            key = (ls[sa], ls[sb], ls[sc], ls[sd])
            if per_class.setdefault(key, 0) >= 60:
                continue
            per_class[key] += 1
            n = hostcheck.hostcheck_shell_quartet(*ob.args(), sa, sb, sc, sd, 4, far_out)
            min_x, proved = hostcheck.hostcheck_last_min_x(), hostcheck.hostcheck_last_proved_far()
            n_total += 1
            n_proved += proved
            # (the far-field threshold depends on the class: 36 for L <= 2, 40 for L <= 4, else 48)
            x_far = hostcheck.hostcheck_far_threshold(int(ls[sa] + ls[sb] + ls[sc] + ls[sd]))
            assert not (proved and min_x < x_far), (shift, sa, sb, sc, sd, min_x)
            if min_x < x_far:
                continue
            n_far += 1
            assert hostcheck.hostcheck_shell_quartet(*ob.args(), sa, sb, sc, sd, 1, gen_out) == n
            scale = max(np.abs(gen_out[:n]).max(), 1e-300)
            worst = max(worst, np.abs(far_out[:n] - gen_out[:n]).max() / scale)
            assert np.abs(far_out[:n] - gen_out[:n]).max() < 1e-14
            seen.add((ls[sa], ls[sb], ls[sc], ls[sd]))
    assert len(seen) == 21, sorted(seen)
    assert worst < 2e-13, worst       # relative to the largest integral of the block
    assert n_proved > 0.5 * n_far     # the proof is not vacuous


def _fused_blocks(hostcheck, ob, boys, I_ref, only_far=False):
    """every quartet of FUSED shells (type 3 = sp: functions s, px, py, pz) in class order"""
    ls = np.zeros(ob.n, dtype=np.int32)
    bf = np.zeros(ob.n, dtype=np.int32)
    ns = hostcheck.hostcheck_nshells(*ob.args(), ls, bf)
    nf = lambda t: 4 if t == 3 else (t + 1) * (t + 2) // 2
    pair = lambda a, b: ls[a] * (ls[a] + 1) // 2 + ls[b]
    worst, classes = 0.0, set()
    out = np.zeros(1296)
    for sa in range(ns):
        for sb in range(ns):
            if ls[sa] < ls[sb]:
                continue
            for sc in range(ns):
                for sd in range(ns):
                    if ls[sc] < ls[sd] or pair(sa, sb) < pair(sc, sd):
                        continue
                    n = hostcheck.hostcheck_shell_quartet(*ob.args(), sa, sb, sc, sd, boys, out)
                    assert n == nf(ls[sa]) * nf(ls[sb]) * nf(ls[sc]) * nf(ls[sd]), (n, ls[sa], ls[sb], ls[sc], ls[sd])
                    if only_far and hostcheck.hostcheck_last_min_x() < 48.0:
                        continue
                    blk = I_ref[bf[sa]:bf[sa] + nf(ls[sa]), bf[sb]:bf[sb] + nf(ls[sb]),
                                bf[sc]:bf[sc] + nf(ls[sc]), bf[sd]:bf[sd] + nf(ls[sd])]
                    worst = max(worst, np.abs(out[:n].reshape(blk.shape) - blk).max())
                    classes.add((ls[sa], ls[sb], ls[sc], ls[sd]))
    return worst, classes, ls


def test_fused_sp_classes_against_oracle(hostcheck, orc, geo, ref_or_restated):
    """The six classes over {s, sp}: an sp shell of STO-3G / 6-31G evaluated as ONE shell whose
    s and p parts share every primitive quartet (generated VRR with per-variant contraction
    weights).  Every block against the oracle's tensor, both Boys flavours, and the far-field
    form on two distant waters."""
    hostcheck.hostcheck_set_fuse(1)
    try:
        z, x = geo.molecule(geo.WATER_CRAWFORD)
        for basis_name in ("6-31G", "STO-3G"):
            ob = orc.make_basis(z, x, basis_name)
            with ref_or_restated():
                I_ref = orc.build_I(ob)
            worst, classes, ls = _fused_blocks(hostcheck, ob, 0, I_ref)
            assert 3 in ls and 1 not in ls  # every p shell was fused with its s partner
            assert classes == {(0, 0, 0, 0), (3, 0, 0, 0), (3, 0, 3, 0), (3, 3, 0, 0), (3, 3, 3, 0),
                               (3, 3, 3, 3)}, sorted(classes)
            assert worst < 1e-12, worst
            I_x = orc.build_I(ob, orc.BOYS_EXACT)
            worst, _, _ = _fused_blocks(hostcheck, ob, 1, I_x)
            assert worst < 1e-13, worst
        # far-field form (rho-free point multipoles) of the fused classes
        z2 = np.concatenate([z, z])
        x2 = np.concatenate([x, x[:, [2, 0, 1]] + np.array([40.0, 9.0, -6.0])])
        ob = orc.make_basis(z2, x2, "6-31G")
        I_x = orc.build_I(ob, orc.BOYS_EXACT)
        worst, classes, _ = _fused_blocks(hostcheck, ob, 4, I_x, only_far=True)
        assert len(classes) == 6 and worst < 1e-14, (worst, sorted(classes))
        # a basis with d shells (or a lone p shell) keeps its shells segmented
        ob = orc.make_basis(z, x, "6-31G*")
        ls = np.zeros(ob.n, dtype=np.int32)
        bf = np.zeros(ob.n, dtype=np.int32)
        ns = hostcheck.hostcheck_nshells(*ob.args(), ls, bf)
        assert 3 not in ls[:ns] and 1 in ls[:ns] and 2 in ls[:ns]
    finally:
        hostcheck.hostcheck_set_fuse(0)


def test_boys_reference_restatement_bitwise_iterations(hostcheck, orc):
    g = golden("fgamma_ref.npz")
    F = np.zeros(9)
    for mode in (2, 0):  # 2 = faithful loops, 0 = fast table-driven path
        worst = 0.0
        for j, x in enumerate(g["x"]):
            hostcheck.hostcheck_boys(mode, 8, float(x), F)
            ref = g["F"][:9, j]
            worst = max(worst, np.abs(F / ref - 1).max())
        # same loops, same iteration counts; only the smooth wrapper differs: the reference forms
        # exp(-x + a ln x - lgamma a) whose argument carries ~|arg| ulp of rounding noise
        assert worst < 1e-13, (mode, worst)


def test_boys_reference_fast_path_equals_faithful_loops(hostcheck):
    """The table-driven reference Boys (converged Boys minus the tabulated truncation error
    of Fgamma, boys_delta.h) must reproduce the faithful series / continued-fraction loops
    everywhere, including right at the iteration-count steps, the branch switch x = m + 3/2
    and the cell boundaries."""
    assert hostcheck.hostcheck_ref_tables_ok() == 1  # at most one step per cell, layout limits ok
    rng = np.random.default_rng(0)
    xs = np.concatenate([
        rng.uniform(0, 70, 60000), np.exp(rng.uniform(np.log(1e-9), np.log(5000), 20000)),
        np.arange(0, 1200) / 16.0, np.nextafter(np.arange(1, 1200) / 16.0, 0),
        np.arange(9) + 1.5, np.nextafter(np.arange(9) + 1.5, 0),
        [0.0, 1e-9, 32, 36, np.nextafter(36, 0), 40, 48, 56, 64, np.nextafter(64, 0), 80, 128, 1e4]])
    xs = np.concatenate([xs, np.arange(0, 2400) / 64.0, np.nextafter(np.arange(1, 2400) / 64.0, 0),
                         (np.arange(9) + 1.5) * (1 + 1e-12), (np.arange(9) + 1.5) * (1 - 1e-12)])
    Fd, Fs = np.zeros(9), np.zeros(9)
    worst_d = 0.0
    for x in xs:
        hostcheck.hostcheck_boys(3, 8, float(x), Fd)   # exact minus tabulated correction (kernels)
        hostcheck.hostcheck_boys(2, 8, float(x), Fs)   # faithful loops
        worst_d = max(worst_d, np.abs(Fd / Fs - 1).max())
    assert worst_d < 2e-14, worst_d
    for L in (0, 2, 4):  # the lower-L instantiations share the tables
        for x in xs[::11]:
            hostcheck.hostcheck_boys(3, L, float(x), Fd)
            hostcheck.hostcheck_boys(2, 8, float(x), Fs)
            assert np.abs(Fd[:L + 1] / Fs[:L + 1] - 1).max() < 2e-14, (L, x)


def test_boys_exact_table(hostcheck, orc):
    L = orc.lib()
    F = np.zeros(9)
    worst = 0.0
    for x in np.concatenate([np.linspace(0, 40, 997), [35.99, 36.0, 36.01, 1e-9, 200.0]]):
        hostcheck.hostcheck_boys(1, 8, float(x), F)
        ref = np.array([L.orc_fgamma_exact(float(m), float(x)) for m in range(9)])
        worst = max(worst, np.abs(F / ref - 1).max())
    assert worst < 5e-15, worst


def test_flop_model_is_generated(rc):
    import json

    fl = json.load(open(os.path.join(ROOT, "rchem_b200", "csrc", "gen", "flops.json")))
    assert len(fl) == 26  # 21 classes over {s,p,d} + 5 more over {s, fused sp}
    # a fused quartet shares the VRR between its s|p variants: (sp sp|sp sp) emits 1210 flops
    # per primitive where the 16 segmented classes it covers emit 1917
    seg = (fl["0000"]["vrr_flops"] + 4 * fl["1000"]["vrr_flops"] + 4 * fl["1010"]["vrr_flops"]
           + 2 * fl["1100"]["vrr_flops"] + 4 * fl["1110"]["vrr_flops"] + fl["1111"]["vrr_flops"])
    assert fl["3333"]["vrr_flops"] < 0.65 * seg and fl["3333"]["out"] == 256
    assert fl["1111"]["hrr_flops"] == 324 and fl["2222"]["hrr_flops"] == 10586
    assert fl["1010"]["vrr_flops"] == 60
    # classes past 100 accumulators are emitted in register-sized parts (one per bra component /
    # bra function); the op-count MODEL stays the monolithic form's, the emitted code recomputes
    # shared intermediates (< 2x)
    assert "vrr_parts" not in fl["2120"] and fl["2120"]["targets"] == 96
    for tag in ("2111", "2121", "2211", "2220", "2221", "2222", "3333"):
        assert fl[tag]["targets"] > 100 and fl[tag]["vrr_parts"] >= 10 and fl[tag]["hrr_parts"] >= 4
        assert fl[tag]["vrr_flops"] < fl[tag]["emitted_vrr_flops"] < 2.0 * fl[tag]["vrr_flops"]
    assert fl["2222"]["vrr_parts"] == 31 and fl["2222"]["hrr_parts"] == 6


# ---- multi-process host path (gloo, world_size 2) -------------------------------------------
_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from rchem_b200 import parallel, geometry as geo
from oracle import oracle as orc
rank, world, _ = parallel.init_distributed("gloo")
z, x = geo.molecule(geo.WATER_CRAWFORD)
ob = orc.make_basis(z, x, "STO-3G")
n = ob.n
D = geo.synthetic_density(n)
I = orc.build_I(ob)
# each rank digests the (mu,nu) rows it owns, like the block-interleaved GPU partition
JK = np.zeros((2, n, n))
for row in range(n * n):
    if parallel.block_owner(row, world) != rank:
        continue
    mu, nu = divmod(row, n)
    JK[0, mu, nu] = (I[mu, nu] * D).sum()
    JK[1, mu, nu] = (I[mu, :, nu, :] * D).sum()
t = torch.from_numpy(JK)
parallel.allreduce_jk(t)
J, K = orc.jk_inmem(I, D)
assert np.abs(t[0].numpy() - J).max() < 1e-13 and np.abs(t[1].numpy() - K).max() < 1e-13
assert sum(parallel.blocks_of_rank(1001, r, world) for r in range(world)) == 1001
dist.barrier()
if rank == 0:
    print("GLOO_OK")
"""


def test_gloo_world_size_2_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                          "29611", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_reference_equals_exact_past_cut(hostcheck):
    """primitive_quartet switches the reference flavour to the exact Boys path once x >=
    ref_exact_from(L) (+0.5 margin): there the two functions agree to < 2e-15 relative."""
    cut = {0: 14.0, 1: 16.0, 2: 18.0, 3: 20.0, 4: 20.0, 5: 22.0, 6: 26.0, 7: 30.0, 8: 36.0}
    Fe, Fs = np.zeros(9), np.zeros(9)
    for m, x0 in cut.items():  # per order (eri_core.h ref_exact_from_order)
        for x in np.concatenate([np.linspace(x0, x0 + 6, 300), np.linspace(x0 + 6, 200, 300)]):
            hostcheck.hostcheck_boys(1, 8, float(x), Fe)
            hostcheck.hostcheck_boys(2, 8, float(x), Fs)
            assert abs(Fs[m] / Fe[m] - 1) < 4e-15, (m, x)


# ---- Basis Set Exchange JSON ingestion ---------------------------------------------
def test_bse_json_ingestion(rc, geo):
    from rchem_b200 import bse

    z, x = geo.molecule(geo.WATER_CRAWFORD)
    path = os.path.join(ROOT, "tests", "golden", "sto-3g.bse.json")
    o, p, off, e, c, nrm = bse.cgtos_from_bse(path, z, x)
    ref = rc.Basis.new(z, x, "STO-3G").export()
    assert np.array_equal(o, ref[0]) and np.array_equal(p, ref[1]) and np.array_equal(off, ref[2])
    # the BSE file carries more digits than the embedded 8-digit table
    assert np.abs(e / ref[3] - 1).max() < 5e-7 and np.abs(c - ref[4]).max() < 5e-8
    b = bse.basis_from_bse(open(path).read(), z, x)
    assert len(b) == 7 and list(b.shells()[0]) == [0, rc.SHELL_SP, 0, 0]
    # a lone d shell and a general contraction: index by position, one CGTO set per row
    custom = {"elements": {"8": {"electron_shells": [
        {"function_type": "gto", "angular_momentum": [2], "exponents": ["0.8"], "coefficients": [["1.0"]]},
        {"function_type": "gto", "angular_momentum": [0], "exponents": ["3.0", "0.5"],
         "coefficients": [["0.4", "0.7"], ["-0.2", "1.0"]]}]}}}
    b = bse.basis_from_bse(custom, [8], [[0.0, 0.0, 0.0]])
    assert len(b) == 6 + 2 and list(b.shells()[0]) == [2, 0, 0]
    with pytest.raises(ValueError):
        bse.basis_from_bse({"elements": {"8": {"electron_shells": [
            {"function_type": "gto_spherical", "angular_momentum": [2], "exponents": ["0.8"],
             "coefficients": [["1.0"]]}]}}}, [8], [[0, 0, 0]])
    with pytest.raises(KeyError):
        bse.basis_from_bse(custom, [1], [[0, 0, 0]])


def test_k_row_fixed_point_bound_is_rigorous(hostcheck, orc, geo):
    """The block kernel accumulates its shared-memory K rows in 64-bit fixed point scaled from
    the bound  sum |terms of one K element| <= 16 Q_ab max|D| max_X sum_Y n_Y Q_XY
    (eri_kernel.cuh krow_add, engine.cu kbound).  Check the inequality -- without the safety
    factor 16 it must already hold with the factor n_B of the bra's second shell -- on every
    (bra shell pair, bra function, ket function) of water / 6-31G* with oracle integrals."""
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    ob = orc.make_basis(z, x, "6-31G*")
    n = ob.n
    I = np.abs(orc.build_I(ob, orc.BOYS_EXACT))
    ls = np.zeros(n, dtype=np.int32)
    bf = np.zeros(n, dtype=np.int32)
    ns = hostcheck.hostcheck_nshells(*ob.args(), ls, bf)
    ls, bf = ls[:ns], bf[:ns]
    nc = (ls + 1) * (ls + 2) // 2
    sh_of = np.repeat(np.arange(ns), nc)
    qf = np.sqrt(np.einsum("ijij->ij", I))                      # per function pair
    Q = np.zeros((ns, ns))
    for s in range(ns):
        for t in range(ns):
            Q[s, t] = qf[np.ix_(sh_of == s, sh_of == t)].max()  # shell-pair Schwarz bound
    assert (I <= np.einsum("ij,kl->ijkl", qf, qf) * (1 + 1e-10) + 1e-15).all()  # Cauchy-Schwarz
    Rmax = (Q * nc[None, :]).sum(axis=1).max()
    D = np.abs(geo.synthetic_density(n) * 50.0)
    Dmax = D.max()
    worst = 0.0
    for A in range(ns):
        for B in range(ns):
            fa, fb = np.where(sh_of == A)[0], np.where(sh_of == B)[0]
            # K[a, x] receives sum_{b in B} sum_l (ab|xl) D[b,l] from bra pair (A,B), all kets
            lhs = np.einsum("abxl,bl->ax", I[np.ix_(fa, fb, np.arange(n), np.arange(n))], D[fb, :])
            bound = Q[A, B] * Dmax * Rmax * len(fb)
            worst = max(worst, lhs.max() / max(bound, 1e-300))
            assert lhs.max() <= bound * (1 + 1e-9) + 1e-14, (A, B, lhs.max(), bound)
    assert 1e-4 < worst <= 1.0  # the bound is neither violated nor absurdly loose


def test_split_word_fixed_point_accumulation_is_exact():
    """krow_add (eri_kernel.cuh) adds a signed 64-bit fixed-point value as two 32-bit atomics:
    low word first (the returned old value gives the carry), then high word + carry.  Restated
    here word for word: whatever the order of the additions, (hi, lo) is the exact 64-bit sum,
    and converting back with the inverse scale recovers the fp64 sum to the resolution."""
    M32, M64 = 1 << 32, 1 << 64
    rng = np.random.default_rng(3)
    for trial in range(50):
        bound = 10.0 ** rng.uniform(-6, 4)
        e = min(61 - (int(np.floor(np.log2(bound))) + 1), 900)   # scale 2^e, bound < 2^(61-e)
        vals = rng.standard_normal(400) * bound / 400.0 * rng.choice([1.0, 1e-6, 1e-12], 400)
        xs = [int(np.rint(v * 2.0 ** e)) for v in vals]
        for order in (np.arange(400), rng.permutation(400)):
            lo = hi = 0
            for i in order:
                xu = xs[i] % M64
                xl, xh = xu % M32, xu >> 32
                old, lo = lo, (lo + xl) % M32
                xh = (xh + (1 if (old + xl) % M32 < old else 0)) % M32
                if xh:
                    hi = (hi + xh) % M32
            acc = (hi << 32) | lo
            acc -= M64 if acc >= 1 << 63 else 0
            assert acc == sum(xs)
            assert abs(acc * 2.0 ** -e - vals.sum()) <= 400 * 2.0 ** -(e + 1) + 1e-18 * bound
