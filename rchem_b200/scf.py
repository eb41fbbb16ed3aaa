"""Host-side mirror of the reference's RHF driver (src/bin/rchem.rs:10-127), the caller of the
hot path: core guess, symmetric orthogonalisation, F = H + 2J - K with J/K from the GPU,
E = sum((H + F) * D), stop at |dE| < 1e-11.  Dense N x N linear algebra stays on the host
(numpy / LAPACK), as it does in the reference (ndarray-linalg)."""
import numpy as np

from . import JK_direct, S, T, V


def build_density(C, nocc):
    """rchem.rs:101-104"""
    return C[:, :nocc] @ C[:, :nocc].T


def calc_elec_energy(D, H, F):
    """rchem.rs:106-108"""
    return float(((H + F) * D).sum())


def build_fock(F, J, K, D, H, basis_set):
    """rchem.rs:110-127"""
    JK_direct(J, K, basis_set, D)
    F[...] = 2.0 * J - K + H


def rhf(basis_set, atomnos, atomcoords, nocc=5, thresh_e=1.0e-11, max_iterations=1024,
        verbose=False):
    """Returns (E_elec, iterations, C, D) like the loop of rchem.rs:40-89 (nocc = 5 is the
    reference's hard-coded water value)."""
    Smat = S(basis_set)
    w, U = np.linalg.eigh(Smat)
    X = U @ np.diag(w ** -0.5) @ U.T  # symm_orthog, rchem.rs:44-52
    H = T(basis_set) + V(basis_set, atomcoords, atomnos)
    n = basis_set.nbf

    def solve(Fmat):
        _, Cp = np.linalg.eigh(X.T @ Fmat @ X)
        return X @ Cp

    C = solve(H)
    D = build_density(C, nocc)
    e_new = calc_elec_energy(D, H, H)
    J, K, F = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    iteration = 0
    while iteration < max_iterations:
        build_fock(F, J, K, np.ascontiguousarray((D + D.T) * 0.5), H, basis_set)
        C = solve(F)
        D = build_density(C, nocc)
        e_old, e_new = e_new, calc_elec_energy(D, H, F)
        if verbose:
            print(f"{iteration:4d} {e_new:20.12f} {e_new - e_old:20.12f}")
        if abs(e_new - e_old) < thresh_e:
            break
        iteration += 1
    return e_new, iteration, C, D
