"""CPU tests that PIN the oracle: against the reference's own golden values, against the
committed fixtures generated from the unmodified reference C sources, and (when
oracle/_ref is present) against that library directly."""
import numpy as np
import pytest

from conftest import golden

RA, RB, RC_, RD = [1.0, 0.0, 1.0], [0.0, 1.0, 2.0], [0.0, 0.0, 3.0], [0.0, 0.0, 4.0]
AL, NR = [1.1, 1.2, 1.3, 1.4], [1.0] * 4


def test_reference_golden_tho66(orc):
    # tho66.rs:288-321 -- the only reference test on the executed ERI kernel, thresh 1e-12
    v = orc.coulomb_repulsion(RA, RB, RC_, RD, [0] * 12, AL, NR)
    assert abs(v - 0.08608517834596989) < 1.0e-12


@pytest.mark.parametrize("powers,value", [
    ([0] * 12, 0.08608517834596989),
    ([1] + [0] * 11, -0.046888060557776973),
    ([2, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 0], 1.71817807954e-05),
])
def test_reference_golden_os86_exact_twin(orc, powers, value):
    # os86.rs:882-937 -- exact-Boys known answers, thresh 1e-15; third value is given to 12 digits
    v = orc.coulomb_repulsion(RA, RB, RC_, RD, powers, AL, NR, orc.BOYS_EXACT)
    assert abs(v - value) < (1.0e-15 if abs(value) > 1e-3 else 1.0e-16)
    # the reference's inexact Boys is within 1e-12 of them (SURVEY section 4 probe)
    assert abs(orc.coulomb_repulsion(RA, RB, RC_, RD, powers, AL, NR) - value) < 1.0e-12


def test_fgamma_matches_reference_fixture(orc):
    g = golden("fgamma_ref.npz")
    L = orc.lib()
    mine = np.array([[L.orc_fgamma_ref(m, x) for x in g["x"]] for m in g["m"]])
    assert np.array_equal(mine, g["F"])  # bit-exact restatement


def test_fgamma_is_inexact_like_the_reference(orc):
    # Fgamma(0,0) = 0.99999999667 (clamp), not 1
    L = orc.lib()
    assert abs(L.orc_fgamma_ref(0.0, 0.0) - 0.99999999667) < 1e-11
    assert abs(L.orc_fgamma_exact(0.0, 0.0) - 1.0) < 1e-16


def test_exact_boys_against_mpmath(orc):
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.dps = 40
    L = orc.lib()
    for m in (0, 1, 4, 8, 12):
        for x in (0.0, 1e-9, 0.3, 2.5, 11.0, 35.9, 36.0, 80.0, 500.0):
            exact = mpmath.quad(lambda t: t ** (2 * m) * mpmath.e ** (-x * t * t), [0, 1])
            got = L.orc_fgamma_exact(float(m), x)
            assert abs(got - float(exact)) <= 4e-16 * float(exact) + 1e-300, (m, x)


def test_primitives_match_reference_fixture(orc):
    g = golden("primitives_ref.npz")
    for t in range(len(g["tho"])):
        c = g["centres"][t]
        v = orc.coulomb_repulsion(c[0:3], c[3:6], c[6:9], c[9:12], g["powers"][t], g["alphas"][t],
                                  g["norms"][t])
        assert v == g["tho"][t], t  # bit-exact restatement of cints.c:72-115
        # and the reference's HGP kernel (chgp.c) agrees with its THO kernel
        assert abs(g["hgp"][t] - g["tho"][t]) < 1e-13 * max(1.0, abs(g["tho"][t]))


def test_restatement_vs_reference_library_directly(orc):
    if orc.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no reference checkout)")
    rng = np.random.default_rng(3)
    for _ in range(200):
        c = rng.uniform(-2, 2, 12)
        al = np.exp(rng.uniform(-2, 4, 4))
        nr = rng.uniform(0.5, 2, 4)
        pw = rng.integers(0, 3, 12).astype(np.int32)
        a = orc.coulomb_repulsion(c[0:3], c[3:6], c[6:9], c[9:12], pw, al, nr)
        b = orc.ref_coulomb_repulsion(c[0:3], c[3:6], c[6:9], c[9:12], pw, al, nr)
        assert a == b


@pytest.mark.parametrize("tag", ["water_crawford", "water"])
def test_water_sto3g_tensor_and_jk_fixture(orc, geo, tag):
    g = golden(f"{tag}_sto3g.npz")
    b = orc.make_basis(g["atomnos"], g["coords"], "STO-3G")
    assert b.n == 7
    I = orc.build_I(b)
    assert np.array_equal(I, g["I"])
    J, K = orc.jk_direct(b, g["D"])
    assert np.abs(J - g["J"]).max() < 1e-15 and np.abs(K - g["K"]).max() < 1e-15
    # JK_inmem on the stored tensor == JK_direct (basis.rs:462-484 vs 383-428)
    J2, K2 = orc.jk_inmem(I, g["D"])
    assert np.abs(J2 - J).max() < 1e-14 and np.abs(K2 - K).max() < 1e-14
    # (ij|kl) = (kl|ij) to rounding (5.6e-16)
    assert np.abs(I - I.transpose(2, 3, 0, 1)).max() < 1e-14


def test_survey_observed_values(orc, geo):
    # SURVEY 8(c) "oracle outputs observed here", 1-based (ij|kl)
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    I = orc.build_I(orc.make_basis(z, x, "STO-3G"))
    for (i, j, k, l), v in {(1, 1, 1, 1): 4.785065547112582, (2, 1, 1, 1): 0.741380377866034,
                            (2, 2, 1, 1): 1.118946911199975, (6, 6, 6, 6): 0.774605927224246,
                            (7, 6, 7, 6): 0.017861326525483, (6, 3, 6, 3): 0.075103218181942}.items():
        assert abs(I[i - 1, j - 1, k - 1, l - 1] - v) < 5e-15
    # Crawford's public values agree to ~1e-7 only (reference Boys error + table digits)
    assert abs(I[0, 0, 0, 0] - 4.785065404705506) < 1e-6


def test_d_shell_fixture(orc):
    g = golden("water_crawford_631gs.npz")
    b = orc.make_basis(g["atomnos"], g["coords"], "6-31G*")
    assert b.n == 19
    vals = orc.eval_quartets(b, g["idx"][:1500])
    # the fixture holds build_I's canonical-order values; (ab|cd) vs (ba|cd) differ by rounding
    assert np.abs(vals - g["vals"][:1500]).max() < 1e-14


def test_canonical_index(orc):
    g = golden("ijkl_ref.npz")
    L = orc.lib()
    got = np.array([L.orc_ijkl2intindex(*map(int, q)) for q in g["ijkl"]])
    assert np.array_equal(got, g["index"])
    # 64-bit: no overflow past N=361
    assert L.orc_ijkl2intindex(1247, 1247, 1247, 1247) == (1247 * 1248 // 2 + 1247) * (1247 * 1248 // 2 + 1248) // 2 + 1247 * 1248 // 2 + 1247


def test_component_order_shell_rs(orc):
    # shell.rs:44-61
    assert orc.ijk_list(3).tolist() == [[3, 0, 0], [2, 1, 0], [2, 0, 1], [1, 2, 0], [1, 1, 1],
                                        [1, 0, 2], [0, 3, 0], [0, 2, 1], [0, 1, 2], [0, 0, 3]]
    assert orc.ijk_list(2).tolist() == [[2, 0, 0], [1, 1, 0], [1, 0, 1], [0, 2, 0], [0, 1, 1], [0, 0, 2]]


def test_normalization_formula(orc):
    # basis.rs:140-149 for an s and a d_xy primitive
    L = orc.lib()
    a = 0.8
    s = L.orc_normalization(np.array([0, 0, 0], dtype=np.int32), a)
    assert abs(s - (2 * a / np.pi) ** 0.75) < 1e-15
    dxy = L.orc_normalization(np.array([1, 1, 0], dtype=np.int32), a)
    dxx = L.orc_normalization(np.array([2, 0, 0], dtype=np.int32), a)
    assert abs(dxy / dxx - np.sqrt(3.0)) < 1e-15


def test_quartet_list_builder(orc):
    Qb = np.array([3.0, 1.0, 0.5, 1e-4, 1e-7])
    out = orc.quartet_list(Qb, Qb, True, 1e-6)
    expect = [(p, q) for p in range(5) for q in range(p + 1) if Qb[p] * Qb[q] >= 1e-6]
    assert [tuple(r) for r in out.tolist()] == expect
    out = orc.quartet_list(Qb, Qb[:3], False, 0.0)
    assert len(out) == 15


# ---- one-electron integrals (the step before the hot path; exact-Boys os86 semantics) ----------
def test_one_electron_golden_os86(orc):
    # os86.rs:939-1004: get_overlap / get_kinetic / get_nuclear known answers
    L = orc.lib()
    ra, rb, rc = np.zeros(3), np.array([0.5, 0.8, -0.2]), np.array([0.5, 0.8, 0.2])
    pw = lambda *a: np.array(a, dtype=np.int32)
    for p, v in (((0, 0, 0, 0, 0, 0), 0.20373275913014607), ((1, 0, 0, 0, 0, 0), 0.062005622343957505),
                 ((1, 1, 0, 1, 1, 0), -0.00043801221837779696), ((2, 1, 0, 1, 1, 0), -0.0002385994651113168)):
        assert abs(L.orc_overlap(1.8, 2.8, ra, rb, pw(*p)) - v) < 1e-16
    for p, v in (((0, 0, 0, 0, 0, 0), 0.3652714583525358), ((1, 0, 0, 0, 0, 0), 0.2514265587836556),
                 ((2, 2, 2, 2, 2, 2), -7.40057384314e-05)):
        assert abs(L.orc_kinetic(1.8, 2.0, ra, rb, pw(*p)) - v) < (3e-16 if abs(v) > 1e-3 else 1e-16)
    for p, v in (((0, 0, 0, 0, 0, 0), -0.49742209545104593), ((1, 0, 0, 0, 0, 0), -0.15987439458254471),
                 ((2, 2, 2, 0, 0, 0), -0.003801373531942607), ((1, 1, 1, 1, 1, 1), 8.8415484347060993e-5)):
        assert abs(L.orc_nuclear(1.8, 2.0, ra, rb, rc, pw(*p)) - v) < 1e-15


def test_one_electron_matrices_water(orc, geo):
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = orc.make_basis(z, x, "STO-3G")
    S = orc.one_electron(b, "S")
    # contraction coefficients are not renormalised: diag(S) = 1 +- 3e-8 (SURVEY 8a note)
    assert np.abs(np.diag(S) - 1).max() < 1e-7 and np.abs(S - S.T).max() == 0
    T, V = orc.one_electron(b, "T"), orc.one_electron(b, "V", z, x)
    # Crawford's published core Hamiltonian elements for this geometry/basis
    H = T + V
    assert abs(H[0, 0] - (-32.5773954)) < 2e-6 and abs(H[1, 0] - (-7.5788328)) < 2e-6


def test_sampled_parity_helper_equals_full_oracle_jk(orc, geo):
    """oracle/parity.py (the checker of the BASELINE-size GPU tests and of bench.py's parity
    block): element rows rebuilt under a Schwarz list equal the full oracle J/K when nothing is
    screened, and drop exactly the screened quartets otherwise."""
    from oracle import parity

    z, x = geo.water_cluster(2)
    ob = orc.make_basis(z, x, "STO-3G")
    n = ob.n
    D = geo.synthetic_density(n)
    I = orc.build_I(ob)
    J, K = orc.jk_inmem(I, D)
    l = np.array([0, 0, 1, 0, 0] * 2, dtype=np.int32)
    first = np.array([0, 1, 2, 5, 6, 7, 8, 9, 12, 13], dtype=np.int32)
    ns = len(l)
    ii, jj = np.tril_indices(ns)
    swap = l[ii] < l[jj]
    sa, sb = np.where(swap, jj, ii), np.where(swap, ii, jj)
    ncart = lambda m: (m + 1) * (m + 2) // 2
    Q = np.array([np.sqrt(max(abs(I[first[a] + i, first[b] + j, first[a] + i, first[b] + j])
                              for i in range(ncart(l[a])) for j in range(ncart(l[b]))))
                  for a, b in zip(sa, sb)])
    fn_shell, Qm = parity.shell_maps(l, first, sa, sb, Q)
    assert list(fn_shell[:7]) == [0, 1, 2, 2, 2, 3, 4] and np.array_equal(Qm, Qm.T)
    els = parity.pick_elements(l, first, sa, sb, Q, 10)
    assert len(els) == 10 and els == parity.pick_elements(l, first, sa, sb, Q, 10)
    Jr, Kr = parity.jk_elements(orc, ob, D, els, fn_shell, Qm, 0.0)
    mu, nu = np.array(els).T
    assert np.abs(Jr - J[mu, nu]).max() < 1e-13 and np.abs(Kr - K[mu, nu]).max() < 1e-13
    # with screening: identical to masking the tensor by the same rule
    tau = 1e-3
    Qf = Qm[np.ix_(fn_shell, fn_shell)]
    mask = Qf[:, :, None, None] * Qf[None, None, :, :] >= tau
    Js = np.einsum("mnls,ls->mn", I * mask, D)
    Ks = np.einsum("mlns,ls->mn", I * mask, D)
    Jr, Kr = parity.jk_elements(orc, ob, D, els, fn_shell, Qm, tau)
    assert np.abs(Jr - Js[mu, nu]).max() < 1e-13 and np.abs(Kr - Ks[mu, nu]).max() < 1e-13
    assert np.abs(Js - J).max() > 1e-6  # (the screening removed something)
