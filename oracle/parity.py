"""Sampled J/K parity against the oracle.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
checker legs); nothing under rchem_b200/ imports this.

At the BASELINE sizes (N = 416 ... 1248) the oracle cannot build a whole J or K, but one
ELEMENT is N^2 contracted integrals: J[mu,nu] = sum_{la,si} (mu nu|la si) D[la,si] and
K[mu,nu] = sum_{la,si} (mu la|nu si) D[la,si]  (basis.rs:406-417).  The rows are rebuilt from
the oracle's integrals (the reference's own coulomb_repulsion when oracle/_ref is present)
with exactly the product's Schwarz screening applied -- a shell quartet is kept iff
Q_bra * Q_ket >= tau on the Q array the library itself reports (rchem_schwarz) -- so the
comparison is at the 1e-12 parity tolerance, not at a screening-error bound.
"""
import numpy as np


def nfun(l):
    """functions of a shell as rchem_basis_shells reports it: l = -1 is a fused sp shell
    (s, px, py, pz)"""
    l = int(l)
    return 4 if l == -1 else (l + 1) * (l + 2) // 2


def shell_maps(shell_l, shell_first, sa, sb, Q):
    """(function -> shell, dense symmetric Q matrix over shells)"""
    ns = len(shell_l)
    nbf = int(shell_first[-1] + nfun(shell_l[-1]))
    fn_shell = np.zeros(nbf, dtype=np.int64)
    for s in range(ns):
        fn_shell[shell_first[s]:shell_first[s] + nfun(shell_l[s])] = s
    Qm = np.zeros((ns, ns))
    Qm[sa, sb] = Q
    Qm[sb, sa] = Q
    return fn_shell, Qm


def pick_elements(shell_l, shell_first, sa, sb, Q, count, seed=7):
    """Seeded (mu, nu) sample.  Half of it walks the shell pairs from the strongest Schwarz
    bound downwards (the strongest pairs of each angular-momentum class first -- the bra pairs
    with the longest ket lists, i.e. the block kernel's --, then quantiles of the Q ordering,
    including a diagonal shell pair); the rest are uniformly random function pairs."""
    rng = np.random.default_rng(seed)
    nbf = int(shell_first[-1] + nfun(shell_l[-1]))
    ncart = nfun
    order = np.argsort(-Q, kind="stable")
    pairs = []
    for want in sorted({(int(shell_l[a]), int(shell_l[b])) for a, b in zip(sa, sb)}):
        for p in order:
            if (shell_l[sa[p]], shell_l[sb[p]]) == want and sa[p] != sb[p]:
                pairs.append(int(p))
                break
    pairs.append(int(next(p for p in order if sa[p] == sb[p])))
    for frac in (0.002, 0.01, 0.03, 0.1, 0.3):
        pairs.append(int(order[int(frac * (len(order) - 1))]))
    out = []
    for p in pairs[:max(1, count // 2)]:
        out.append((int(shell_first[sa[p]] + rng.integers(ncart(shell_l[sa[p]]))),
                    int(shell_first[sb[p]] + rng.integers(ncart(shell_l[sb[p]])))))
    while len(out) < count:
        out.append(tuple(int(v) for v in rng.integers(0, nbf, size=2)))
    return out


def jk_elements(orc, obasis, D, elements, fn_shell, Qm, tau):
    """Oracle values of J[mu,nu] and K[mu,nu] for the listed elements under the product's
    screening rule.  Returns (J_ref, K_ref) arrays aligned with `elements`."""
    n = obasis.n
    la, si = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    la, si = la.reshape(-1), si.reshape(-1)
    Dflat = np.ascontiguousarray(D).reshape(-1)
    Qk_all = Qm[fn_shell[la], fn_shell[si]]
    J_ref, K_ref = [], []
    for mu, nu in elements:
        keep = Qm[fn_shell[mu], fn_shell[nu]] * Qk_all >= tau
        q = np.stack([np.full(keep.sum(), mu), np.full(keep.sum(), nu), la[keep], si[keep]], axis=1)
        J_ref.append(float(np.dot(orc.eval_quartets(obasis, q.astype(np.int32)), Dflat[keep])))
        keep = Qm[fn_shell[mu], fn_shell[la]] * Qm[fn_shell[nu], fn_shell[si]] >= tau
        q = np.stack([np.full(keep.sum(), mu), la[keep], np.full(keep.sum(), nu), si[keep]], axis=1)
        K_ref.append(float(np.dot(orc.eval_quartets(obasis, q.astype(np.int32)), Dflat[keep])))
    return np.array(J_ref), np.array(K_ref)


def sampled_jk_errors(orc, obasis, basis, D, J, K, tau, count=8, seed=7):
    """max |J - J_oracle|, max |K - K_oracle| over a seeded element sample (see pick_elements)."""
    l, first = basis.shells()
    sa, sb, _, Q = basis.schwarz()
    fn_shell, Qm = shell_maps(l, first, sa, sb, Q)
    elements = pick_elements(l, first, sa, sb, Q, count, seed)
    J_ref, K_ref = jk_elements(orc, obasis, D, elements, fn_shell, Qm, tau)
    mu = np.array([e[0] for e in elements])
    nu = np.array([e[1] for e in elements])
    return (float(np.abs(J[mu, nu] - J_ref).max()), float(np.abs(K[mu, nu] - K_ref).max()),
            elements)
