"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the
oracle on identical inputs.  Tolerances: 1e-12 absolute per integral in fp64 (north star);
bit-exact for the screened quartet list."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu
TOL = 1e-12

RA, RB, RC_, RD = [1.0, 0.0, 1.0], [0.0, 1.0, 2.0], [0.0, 0.0, 3.0], [0.0, 0.0, 4.0]


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(rc):
    assert rc.device_count() >= 1, "GPU tests need a CUDA device (there is no CPU fallback)"


# ---- tier 1: the libpyquante2 symbol ---------------------------------------------------------
def test_tier1_reference_golden_values(rc):
    # tho66.rs:288-321 (1e-12) through the pyquante2_coulomb_repulsion-shaped wrapper
    v = rc.coulomb_repulsion(1.1, 1.2, 1.3, 1.4, RA, RB, RC_, RD, 1.0, 1.0, 1.0, 1.0, [0] * 12)
    assert abs(v - 0.08608517834596989) < 1.0e-12
    # os86.rs:882-937 known answers (exact Boys)
    for pw, val in (([0] * 12, 0.08608517834596989), ([1] + [0] * 11, -0.046888060557776973),
                    ([2, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 0], 1.71817807954e-05)):
        got = rc.coulomb_repulsion_batch([RA + RB + RC_ + RD], [[1.0] * 4], [pw],
                                         [[1.1, 1.2, 1.3, 1.4]], rc.BOYS_EXACT)[0]
        assert abs(got - val) < (1e-15 if abs(val) > 1e-3 else 1e-16)


def test_tier1_batch_against_reference_fixture(rc):
    g = golden("primitives_ref.npz")
    got = rc.coulomb_repulsion_batch(g["centres"], g["norms"], g["powers"], g["alphas"])
    assert np.abs(got - g["tho"]).max() < TOL
    assert np.abs(got - g["tho"]).max() < 1e-13 * max(1.0, np.abs(g["tho"]).max())


# ---- dense tensor (build_I) --------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["water_crawford", "water"])
def test_build_I_water_sto3g_fixture(rc, tag):
    # BASELINE.json configs 1 and 2
    g = golden(f"{tag}_sto3g.npz")
    b = rc.Basis.new(g["atomnos"], g["coords"], "STO-3G")
    I = rc.build_I(b)
    assert I.shape == (7, 7, 7, 7)
    assert np.abs(I - g["I"]).max() < TOL
    st = b.stats()
    # 5 segmented shells -> 15 pairs -> 120 canonical quartets (the dense tensor is built by the
    # segmented classes also when the J/K path keeps the sp shell fused)
    assert st["shell_quartets"] == 120 and st["launches"] >= 1


@pytest.mark.parametrize("basis_name", ["STO-3G", "6-31G", "6-31G*"])
def test_build_I_all_classes_vs_oracle(rc, orc, geo, ref_or_restated, basis_name):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = rc.Basis.new(z, x, basis_name)
    ob = orc.make_basis(z, x, basis_name)
    with ref_or_restated():
        I_ref = orc.build_I(ob)
    I = rc.build_I(b)
    assert np.abs(I - I_ref).max() < TOL
    # all 8 index permutations are filled (basis.rs:451-454 plus the bra<->ket swap)
    for perm in ((1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)):
        assert np.array_equal(I, I.transpose(perm))
    # exact-Boys flavour against the exact-Boys twin oracle
    b.set_boys(rc.BOYS_EXACT)
    I_x = rc.build_I(b)
    assert np.abs(I_x - orc.build_I(ob, orc.BOYS_EXACT)).max() < TOL
    assert np.abs(I_x - I_ref).max() < 2e-8  # Boys-limited, not a parity claim


def test_d_shell_fixture(rc):
    g = golden("water_crawford_631gs.npz")
    b = rc.Basis.new(g["atomnos"], g["coords"], "6-31G*")
    I = rc.build_I(b)
    i = g["idx"]
    assert np.abs(I[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] - g["vals"]).max() < TOL
    J, K = np.zeros((19, 19)), np.zeros((19, 19))
    rc.JK_direct(J, K, b, g["D"])
    assert np.abs(J - g["J"]).max() < TOL and np.abs(K - g["K"]).max() < TOL


def test_multicentre_d_sample(rc):
    g = golden("water2_631gs_sample.npz")
    b = rc.Basis.new(g["atomnos"], g["coords"], "6-31G*")
    I = rc.build_I(b)
    i = g["idx"]
    assert np.abs(I[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] - g["vals"]).max() < TOL


def test_build_I_from_explicit_cgtos(rc, orc, geo):
    z, x = geo.water_cluster(2)
    ob = orc.make_basis(z, x, "6-31G")
    b = rc.Basis.from_cgtos(ob.origins, ob.powers, ob.prim_offset, ob.exps, ob.coefs, ob.norms)
    I = rc.build_I(b)
    q = np.random.default_rng(5).integers(0, ob.n, size=(2000, 4)).astype(np.int32)
    assert np.abs(I[q[:, 0], q[:, 1], q[:, 2], q[:, 3]] - orc.eval_quartets(ob, q)).max() < TOL


# ---- J/K --------------------------------------------------------------------------------------
def test_jk_direct_water_fixture_and_oracle(rc, orc):
    g = golden("water_crawford_sto3g.npz")
    b = rc.Basis.new(g["atomnos"], g["coords"], "STO-3G")
    J, K = np.full((7, 7), 99.0), np.full((7, 7), -99.0)  # must be overwritten (basis.rs:389-390)
    rc.JK_direct(J, K, b, g["D"])
    assert np.abs(J - g["J"]).max() < TOL and np.abs(K - g["K"]).max() < TOL
    # the reference's own loop nest (no symmetry, 2 N^4 K^4 primitive calls)
    Jo, Ko = orc.jk_direct(orc.make_basis(g["atomnos"], g["coords"], "STO-3G"), g["D"])
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    # repeated calls give the same answer (buffers are re-zeroed)
    J2, K2 = np.zeros((7, 7)), np.zeros((7, 7))
    rc.JK_direct(J2, K2, b, g["D"])
    assert np.abs(J2 - J).max() < 1e-14 and np.abs(K2 - K).max() < 1e-14


def test_jk_direct_cluster_vs_oracle_tensor(rc, orc, geo, ref_or_restated):
    z, x = geo.water_cluster(3)
    b = rc.Basis.new(z, x, "6-31G")
    ob = orc.make_basis(z, x, "6-31G")
    n = ob.n
    with ref_or_restated():
        I_ref = orc.build_I(ob)
    D = geo.synthetic_density(n)
    Jo, Ko = orc.jk_inmem(I_ref, D)
    J, K = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J, K, b, D)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    # JK_inmem on the GPU-built tensor
    Jg, Kg = rc.JK_inmem(rc.build_I(b), D)
    assert np.abs(Jg - Jo).max() < TOL and np.abs(Kg - Ko).max() < TOL


def test_jk_inmem_fixture(rc):
    g = golden("water_sto3g.npz")
    J, K = rc.JK_inmem(g["I"], g["D"])
    assert np.abs(J - g["J"]).max() < 1e-13 and np.abs(K - g["K"]).max() < 1e-13


def test_asymmetric_density_small(rc, orc, geo):
    # basis.rs:383-428 accepts any D (details: test_gpu_kernels.py::test_asymmetric_density_...)
    z, x = geo.molecule(geo.WATER)
    b = rc.Basis.new(z, x, "STO-3G")
    D = np.arange(49.0).reshape(7, 7) / 49.0
    J, K = np.zeros((7, 7)), np.zeros((7, 7))
    rc.JK_direct(J, K, b, D)
    Jo, Ko = orc.jk_direct(orc.make_basis(z, x, "STO-3G"), D)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL


# ---- screening: Schwarz bounds and the quartet list -----------------------------------------------
def test_schwarz_bounds_and_quartet_list(rc, orc, geo):
    z, x = geo.water_cluster(4)
    b = rc.Basis.new(z, x, "6-31G*")
    ob = orc.make_basis(z, x, "6-31G*")
    sa, sb, batch, Q = b.schwarz()
    l, first = b.shells()
    ns = len(l)
    assert len(Q) == ns * (ns + 1) // 2
    assert np.all(l[sa] >= l[sb])
    # Q = sqrt(max |(ab|ab)|) over the shell pair's components, exact Boys
    rng = np.random.default_rng(2)
    ncart = rc.shell_nfun
    for p in rng.choice(len(Q), size=60, replace=False):
        fa = [first[sa[p]] + i for i in range(ncart(l[sa[p]]))]
        fb = [first[sb[p]] + i for i in range(ncart(l[sb[p]]))]
        quartets = np.array([[i, j, i, j] for i in fa for j in fb], dtype=np.int32)
        ref = np.sqrt(np.abs(orc.eval_quartets(ob, quartets, orc.BOYS_EXACT)).max())
        assert abs(Q[p] - ref) < TOL
    # inside a batch the pairs are sorted by Q, descending
    for bt in np.unique(batch):
        q = Q[batch == bt]
        assert np.all(q[:-1] >= q[1:])
    # the implicit list, materialised on the GPU, equals the CPU builder BIT-EXACTLY on the
    # shared Q array
    for tau in (0.0, 1e-10, 1e-6):
        b.set_schwarz_tau(tau)
        got = b.quartet_list()
        offs = {bt: int(np.flatnonzero(batch == bt)[0]) for bt in np.unique(batch)}
        exp = []
        for bi in sorted(offs):
            for ki in sorted(offs):
                if ki > bi:
                    continue
                pq = orc.quartet_list(Q[batch == bi], Q[batch == ki], bi == ki, tau).astype(np.int64)
                pq[:, 0] += offs[bi]
                pq[:, 1] += offs[ki]
                exp.append(pq)
        exp = np.concatenate(exp)
        assert got.shape == exp.shape and np.array_equal(got, exp)
    npair = len(Q)
    b.set_schwarz_tau(0.0)
    assert len(b.quartet_list()) == npair * (npair + 1) // 2


def test_screened_jk_error_is_bounded(rc, geo):
    z, x = geo.water_cluster(8)
    b = rc.Basis.new(z, x, "6-31G")
    n = b.nbf
    D = geo.synthetic_density(n)
    J0, K0 = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J0, K0, b, D)
    full = b.stats()["shell_quartets"]
    b.set_schwarz_tau(1e-10)
    J1, K1 = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J1, K1, b, D)
    st = b.stats()
    assert st["shell_quartets"] < full and st["shell_quartets_all"] == full
    # every dropped integral is < tau in magnitude; |D| <= ~4/n
    bound = 1e-10 * np.abs(D).sum() * 4
    assert np.abs(J1 - J0).max() < bound and np.abs(K1 - K0).max() < bound


# ---- size-independent properties at a BASELINE size ((H2O)_10 STO-3G, config 3) ----------------------
def test_properties_water10(rc, orc, geo):
    z, x = geo.water_cluster(10)
    b = rc.Basis.new(z, x, "STO-3G")
    ob = orc.make_basis(z, x, "STO-3G")
    n = b.nbf
    assert n == 70
    D1 = geo.synthetic_density(n, seed=1)
    D2 = geo.synthetic_density(n, seed=2)

    def jk(D):
        J, K = np.zeros((n, n)), np.zeros((n, n))
        rc.JK_direct(J, K, b, D)
        return J, K

    J1, K1 = jk(D1)
    J2, K2 = jk(D2)
    J12, K12 = jk(D1 + 2.0 * D2)
    assert np.abs(J12 - (J1 + 2 * J2)).max() < 1e-12 and np.abs(K12 - (K1 + 2 * K2)).max() < 1e-12
    assert np.abs(J1 - J1.T).max() < 1e-13 and np.abs(K1 - K1.T).max() < 1e-13
    # the Coulomb and exchange super-operators are symmetric: <D2|J[D1]> = <D1|J[D2]>
    assert abs((J1 * D2).sum() - (J2 * D1).sum()) < 1e-12
    assert abs((K1 * D2).sum() - (K2 * D1).sum()) < 1e-12
    # spot-check against the oracle: J and K rows rebuilt from oracle integrals
    rng = np.random.default_rng(0)
    for mu, nu in rng.integers(0, n, size=(3, 2)):
        qs = np.array([[mu, nu, la, si] for la in range(n) for si in range(n)], dtype=np.int32)
        assert abs((orc.eval_quartets(ob, qs).reshape(n, n) * D1).sum() - J1[mu, nu]) < TOL
        qs = np.array([[mu, la, nu, si] for la in range(n) for si in range(n)], dtype=np.int32)
        assert abs((orc.eval_quartets(ob, qs).reshape(n, n) * D1).sum() - K1[mu, nu]) < TOL
    # dense tensor (192 MB) vs direct J/K, and a sample of its elements vs the oracle
    I = rc.build_I(b)
    Jm, Km = rc.JK_inmem(I, D1)
    assert np.abs(Jm - J1).max() < 1e-12 and np.abs(Km - K1).max() < 1e-12
    q = rng.integers(0, n, size=(4000, 4)).astype(np.int32)
    assert np.abs(I[q[:, 0], q[:, 1], q[:, 2], q[:, 3]] - orc.eval_quartets(ob, q)).max() < TOL


# ---- device-buffer API and the multi-GPU partition -----------------------------------------------------
def test_device_api_and_rank_partition(rc, geo):
    torch = pytest.importorskip("torch")
    from rchem_b200 import parallel

    z, x = geo.water_cluster(5)
    b = rc.Basis.new(z, x, "6-31G*")
    n = b.nbf
    D = geo.synthetic_density(n)
    J, K = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J, K, b, D)
    dev = torch.device("cuda", 0)
    Dd = torch.from_numpy(D).to(dev)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    total = torch.zeros((2, n, n), dtype=torch.float64, device=dev)
    quartets = 0
    for rank in range(3):  # three "ranks" on one GPU: shares must add up to the whole
        part = torch.empty((2, n, n), dtype=torch.float64, device=dev)
        parallel.jk_direct_distributed(b, Dd, part, rank, 3)
        total += part
        quartets += b.stats()["shell_quartets"]
    torch.cuda.synchronize()
    assert abs(quartets - b.stats()["shell_quartets_all"]) <= 3 * b.stats()["n_tasks"]
    assert np.abs(total[0].cpu().numpy() - J).max() < 1e-12
    assert np.abs(total[1].cpu().numpy() - K).max() < 1e-12
    # dense tensor on a device buffer + JK_inmem on device buffers
    I = torch.empty((n,) * 4, dtype=torch.float64, device=dev)
    b.build_I_device(I.data_ptr())
    JK = torch.empty((2, n, n), dtype=torch.float64, device=dev)
    rc.jk_inmem_device(n, I.data_ptr(), Dd.data_ptr(), JK.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.abs(JK[0].cpu().numpy() - J).max() < 1e-12
    assert np.abs(JK[1].cpu().numpy() - K).max() < 1e-12
    b.use_own_stream()


# ---- the caller of the hot path: the reference's RHF loop (rchem.rs:40-89) -------------------------------
def test_rhf_water_crawford_energy(rc, orc, geo):
    """SCF driven exactly like src/bin/rchem.rs (core guess, symmetric orthogonalisation,
    F = H + 2J - K, nocc = 5, |dE| < 1e-11) with J/K from the GPU path.  One-electron matrices
    come from the reference's own C kernels (test infrastructure).  SURVEY section 6: converged
    E_elec = -82.944446488 with libpyquante2 integrals (Crawford: -82.944446990)."""
    import ctypes as C

    ref = orc.ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not available")
    d, i = C.c_double, C.c_int
    for fn in (ref.overlap, ref.kinetic):
        fn.restype = d
        fn.argtypes = [d, i, i, i, d, d, d, d, i, i, i, d, d, d]
    ref.nuclear_attraction.restype = d
    ref.nuclear_attraction.argtypes = [d, d, d, d, i, i, i, d, d, d, d, d, i, i, i, d, d, d, d]
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    ob = orc.make_basis(z, x, "STO-3G")
    n = ob.n
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for mu in range(n):
        for nu in range(n):
            for p in range(ob.prim_offset[mu], ob.prim_offset[mu + 1]):
                for q in range(ob.prim_offset[nu], ob.prim_offset[nu + 1]):
                    cc = ob.coefs[p] * ob.coefs[q]
                    a = (float(ob.exps[p]), *map(int, ob.powers[mu]), *map(float, ob.origins[mu]))
                    b = (float(ob.exps[q]), *map(int, ob.powers[nu]), *map(float, ob.origins[nu]))
                    nn = ob.norms[p] * ob.norms[q]
                    S[mu, nu] += cc * nn * ref.overlap(*a, *b)
                    T[mu, nu] += cc * nn * ref.kinetic(*a, *b)
                    for zc, xc in zip(z, x):
                        V[mu, nu] += cc * float(zc) * ref.nuclear_attraction(
                            *map(float, ob.origins[mu]), float(ob.norms[p]), *map(int, ob.powers[mu]),
                            float(ob.exps[p]), *map(float, ob.origins[nu]), float(ob.norms[q]),
                            *map(int, ob.powers[nu]), float(ob.exps[q]), *map(float, xc))
    H = T + V
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    basis = rc.Basis.new(z, x, "STO-3G")

    def density(F):
        _, Cp = np.linalg.eigh(X.T @ F @ X)
        Cm = X @ Cp
        return Cm[:, :5] @ Cm[:, :5].T

    D = density(H)
    e_new = ((H + H) * D).sum()
    assert abs(e_new - (-125.842076851)) < 5e-8  # core-guess energy, SURVEY section 6
    J, K = np.zeros((n, n)), np.zeros((n, n))
    for it in range(200):
        rc.JK_direct(J, K, basis, np.ascontiguousarray((D + D.T) / 2))
        F = H + 2.0 * J - K
        D = density(F)
        e_old, e_new = e_new, ((H + F) * D).sum()
        if abs(e_new - e_old) < 1e-11:
            break
    assert it < 100
    assert abs(e_new - (-82.944446488)) < 5e-8
    assert abs(e_new - (-82.944446990)) < 1e-6  # Crawford's published value


# ---- widened rows: one-electron matrices and the RHF driver --------------------
@pytest.mark.parametrize("basis_name", ["STO-3G", "6-31G*"])
def test_one_electron_matrices_vs_oracle(rc, orc, geo, basis_name):
    z, x = geo.water_cluster(2)
    b = rc.Basis.new(z, x, basis_name)
    ob = orc.make_basis(z, x, basis_name)
    for name, got in (("S", rc.S(b)), ("T", rc.T(b)), ("V", rc.V(b, x, z))):
        ref = orc.one_electron(ob, name, z, x)
        assert np.abs(got - ref).max() < 1e-12, name
        assert np.array_equal(got, got.T)


def test_rhf_driver_on_device_integrals(rc, geo):
    """rchem.rs main(): S, T, V and J/K all from the GPU library.  With the exact-Boys
    one-electron integrals of the reference's os86 path and libpyquante2-flavoured J/K the
    converged energy sits within 1e-6 of Crawford's published -82.944446990."""
    from rchem_b200 import scf

    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = rc.Basis.new(z, x, "STO-3G")
    e, its, C, D = scf.rhf(b, z, x)
    assert its < 60
    assert abs(e - (-82.944446990)) < 1e-6
    b.set_boys(rc.BOYS_EXACT)
    e2, _, _, _ = scf.rhf(b, z, x)
    assert abs(e2 - e) < 1e-6 and abs(e2 - (-82.944446990)) < 1e-6


# ---- edge cases: tiny, ragged and degenerate inputs ---------------------------------------------------
def test_single_function_and_single_shell(rc, orc):
    # one H atom, STO-3G: one s function, one shell, one pair, one quartet
    z, x = np.array([1], dtype=np.uint64), np.zeros((1, 3))
    b = rc.Basis.new(z, x, "STO-3G")
    ob = orc.make_basis(z, x, "STO-3G")
    I = rc.build_I(b)
    assert I.shape == (1, 1, 1, 1) and abs(I[0, 0, 0, 0] - orc.contracted_eri(ob, 0, 0, 0, 0)) < TOL
    J, K = np.zeros((1, 1)), np.zeros((1, 1))
    rc.JK_direct(J, K, b, np.array([[0.7]]))
    assert abs(J[0, 0] - 0.7 * I[0, 0, 0, 0]) < TOL and abs(K[0, 0] - 0.7 * I[0, 0, 0, 0]) < TOL
    # one O atom, 6-31G*: every shell on one centre (all Boys arguments are 0 -> the 1e-8 clamp)
    z, x = np.array([8], dtype=np.uint64), np.array([[0.3, -0.2, 0.1]])
    b = rc.Basis.new(z, x, "6-31G*")
    ob = orc.make_basis(z, x, "6-31G*")
    I = rc.build_I(b)
    q = np.random.default_rng(9).integers(0, ob.n, size=(1500, 4)).astype(np.int32)
    assert np.abs(I[q[:, 0], q[:, 1], q[:, 2], q[:, 3]] - orc.eval_quartets(ob, q)).max() < TOL


def test_ragged_custom_basis(rc, orc):
    """Shells with ragged contraction lengths (1..5 primitives), random exponents, coefficients
    of both signs and s/p/d mixed on four centres, passed as explicit CGTOs."""
    rng = np.random.default_rng(123)
    centres = rng.uniform(-1.8, 1.8, size=(4, 3))
    origins, powers, off, exps, coefs, norms = [], [], [0], [], [], []
    L = orc.lib()
    for c, (l, k) in zip([0, 0, 1, 1, 2, 3, 3, 2], [(0, 5), (1, 2), (0, 1), (2, 1), (1, 3), (0, 4), (2, 2), (0, 2)]):
        e = np.exp(rng.uniform(np.log(0.2), np.log(30.0), size=k))
        cf = rng.uniform(-1.0, 1.0, size=k)
        for pw in orc.ijk_list(l):
            origins.append(centres[c]); powers.append(pw)
            for ee, cc in zip(e, cf):
                exps.append(ee); coefs.append(cc)
                norms.append(L.orc_normalization(np.ascontiguousarray(pw), ee))
            off.append(len(exps))
    ob = orc.FlatBasis(origins, powers, off, exps, coefs, norms)
    b = rc.Basis.from_cgtos(ob.origins, ob.powers, ob.prim_offset, ob.exps, ob.coefs, ob.norms)
    n = ob.n
    assert n == 1 + 3 + 1 + 6 + 3 + 1 + 6 + 1
    I = rc.build_I(b)
    I_ref = orc.build_I(ob)
    assert np.abs(I - I_ref).max() < TOL * max(1.0, np.abs(I_ref).max())
    D = np.random.default_rng(5).standard_normal((n, n)); D = (D + D.T) / 2
    J, K = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J, K, b, D)
    Jo, Ko = orc.jk_inmem(I_ref, D)
    scale = max(1.0, np.abs(Jo).max(), np.abs(Ko).max())
    assert np.abs(J - Jo).max() < TOL * scale and np.abs(K - Ko).max() < TOL * scale


def test_everything_screened_and_zero_density(rc, geo):
    z, x = geo.water_cluster(2)
    b = rc.Basis.new(z, x, "STO-3G")
    n = b.nbf
    D = geo.synthetic_density(n)
    b.set_schwarz_tau(1e30)  # nothing survives
    J, K = np.ones((n, n)), np.ones((n, n))
    rc.JK_direct(J, K, b, D)
    assert not J.any() and not K.any() and b.stats()["shell_quartets"] == 0
    assert not rc.build_I(b).any()
    b.set_schwarz_tau(0.0)
    rc.JK_direct(J, K, b, np.zeros((n, n)))
    assert not J.any() and not K.any()


def test_prim_eps_option_does_not_change_results(rc, orc, geo):
    z, x = geo.water_cluster(3)
    n = None
    out = []
    for eps in (1e-20, 0.0):
        b = rc.Basis.new(z, x, "6-31G")
        b.set_prim_eps(eps)
        n = b.nbf
        D = geo.synthetic_density(n)
        J, K = np.zeros((n, n)), np.zeros((n, n))
        rc.JK_direct(J, K, b, D)
        out.append((J, K, b.stats()["prim_quartets"]))
    assert out[0][2] < out[1][2]  # primitive pairs were dropped ...
    assert np.abs(out[0][0] - out[1][0]).max() < 1e-13 and np.abs(out[0][1] - out[1][1]).max() < 1e-13


# ---- far-field scheduling (primitive_quartet_far through the block and light kernels) -----------
def test_far_field_pair_of_waters_vs_oracle(rc, orc, geo, ref_or_restated):
    """Two waters 25 bohr apart, 6-31G*: most inter-molecular quartets are PROVED far-field and
    go through the point-multipole form; J/K must still match the reference to 1e-12."""
    z1, x1 = geo.molecule(geo.WATER_CRAWFORD)
    z = np.concatenate([z1, z1])
    x = np.concatenate([x1, x1[:, [2, 0, 1]] + np.array([25.0, 3.0, -4.0])])
    b = rc.Basis.new(z, x, "6-31G*")
    ob = orc.make_basis(z, x, "6-31G*")
    n = ob.n
    D = geo.synthetic_density(n)
    with ref_or_restated():
        Jo, Ko = orc.jk_inmem(orc.build_I(ob), D)
    J, K = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J, K, b, D)
    assert np.abs(J - Jo).max() < TOL and np.abs(K - Ko).max() < TOL
    b.set_far_sched(False)  # the general code only
    J0, K0 = np.zeros((n, n)), np.zeros((n, n))
    rc.JK_direct(J0, K0, b, D)
    assert np.abs(J0 - Jo).max() < TOL and np.abs(K0 - Ko).max() < TOL
    assert np.abs(J - J0).max() < 1e-13 and np.abs(K - K0).max() < 1e-13


@pytest.mark.parametrize("nw,basis_name,tau", [(27, "6-31G", 0.0), (27, "6-31G", 1e-10),
                                               (12, "6-31G*", 0.0), (27, "STO-3G", 1e-10)])
def test_far_scheduling_equals_general_code(rc, geo, nw, basis_name, tau):
    """Clusters large enough for the block kernel (heavy bra pairs) and the light kernel to
    sort their kets into [far | grid | corrected]: J/K with the far-field routing must equal
    J/K with every quartet through the general code (itself checked against the oracle
    above), in both Boys flavours."""
    z, x = geo.water_cluster(nw)
    b = rc.Basis.new(z, x, basis_name)
    b.set_schwarz_tau(tau)
    n = b.nbf
    D = geo.synthetic_density(n)
    for boys in (rc.BOYS_REFERENCE, rc.BOYS_EXACT):
        b.set_boys(boys)
        out = []
        for far in (True, False):
            b.set_far_sched(far)
            J, K = np.zeros((n, n)), np.zeros((n, n))
            rc.JK_direct(J, K, b, D)
            out.append((J, K))
        assert np.abs(out[0][0] - out[1][0]).max() < 1e-12, (boys, np.abs(out[0][0] - out[1][0]).max())
        assert np.abs(out[0][1] - out[1][1]).max() < 1e-12, (boys, np.abs(out[0][1] - out[1][1]).max())
        assert np.abs(out[0][0]).max() > 1e-3
