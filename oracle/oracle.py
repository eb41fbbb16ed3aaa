"""ctypes front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under rchem_b200/ does.

It wraps

* ``liberi_oracle.so``  -- our C restatement (oracle/eri_oracle.c), and
* ``_ref/libpyquante2_ref.so`` -- the UNMODIFIED reference C sources (cints.c, chgp.c)
  compiled by oracle/Makefile, when present.  ``use_reference_kernel(True)`` makes the
  restated basis.rs loop nests call the reference's own ``coulomb_repulsion``.

and carries an independent Python statement of ``Basis::new`` (basis.rs:182-211) with the
same embedded basis tables the product embeds, so the product's data model can be
cross-checked against it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BOYS_REFERENCE, BOYS_EXACT = 0, 1

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the oracle (and oracle/_ref when the reference checkout is present)."""
    so = os.path.join(HERE, "liberi_oracle.so")
    ref = os.path.join(HERE, "_ref", "libpyquante2_ref.so")
    src = os.path.join(HERE, "eri_oracle.c")
    stale = not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale or (os.path.isdir("/root/reference/libpyquante2") and not os.path.exists(ref)):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return so


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_fgamma_ref.restype = C.c_double
        L.orc_fgamma_ref.argtypes = [C.c_double, C.c_double]
        L.orc_fgamma_exact.restype = C.c_double
        L.orc_fgamma_exact.argtypes = [C.c_double, C.c_double]
        L.orc_coulomb_repulsion.restype = C.c_double
        L.orc_coulomb_repulsion.argtypes = [_dp, _dp, _dp, _dp, _ip, _dp, _dp, C.c_int]
        L.orc_normalization.restype = C.c_double
        L.orc_normalization.argtypes = [_ip, C.c_double]
        L.orc_ijk_list.restype = C.c_int
        L.orc_ijk_list.argtypes = [C.c_int, _ip]
        L.orc_ijkl2intindex.restype = C.c_int64
        L.orc_ijkl2intindex.argtypes = [C.c_int64] * 4
        basis_args = [C.c_int, _dp, _ip, _ip, _dp, _dp, _dp]
        L.orc_contracted_eri.restype = C.c_double
        L.orc_contracted_eri.argtypes = basis_args + [C.c_int] * 5
        L.orc_build_I.restype = None
        L.orc_build_I.argtypes = basis_args + [_dp, C.c_int]
        L.orc_jk_direct.restype = None
        L.orc_jk_direct.argtypes = basis_args + [_dp, _dp, _dp, C.c_int]
        L.orc_jk_inmem.restype = None
        L.orc_jk_inmem.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.orc_eval_quartets.restype = C.c_double
        L.orc_eval_quartets.argtypes = basis_args + [_ip, C.c_int64, C.c_void_p, C.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        L.orc_get_max_threads.argtypes = []
        L.orc_set_reference_kernel.restype = None
        L.orc_set_reference_kernel.argtypes = [C.c_void_p]
        L.orc_one_electron.restype = None
        L.orc_one_electron.argtypes = basis_args + [C.c_int, C.c_int, _dp, _dp, _dp]
        for f in (L.orc_overlap, L.orc_kinetic):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_double, _dp, _dp, _ip]
        L.orc_nuclear.restype = C.c_double
        L.orc_nuclear.argtypes = [C.c_double, C.c_double, _dp, _dp, _dp, _ip]
        L.orc_quartet_list.restype = C.c_int64
        L.orc_quartet_list.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double,
                                       C.c_void_p, C.c_int64]
    return _lib


def ref_lib():
    """The unmodified reference C library, or None when it was not built/shipped."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libpyquante2_ref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        d, i = C.c_double, C.c_int
        centre = [d, d, d, d, i, i, i, d]
        _ref.coulomb_repulsion.restype = d
        _ref.coulomb_repulsion.argtypes = centre * 4  # cints.h:23-30
        _ref.hrr.restype = d
        _ref.hrr.argtypes = centre * 4  # chgp.h
        _ref.Fgamma.restype = d
        _ref.Fgamma.argtypes = [d, d]
        _ref.ijkl2intindex.restype = i
        _ref.ijkl2intindex.argtypes = [i] * 4
    return _ref


def set_num_threads(n):
    """OpenMP threads of the oracle's loops (torchrun exports OMP_NUM_THREADS=1); returns the
    thread count actually in effect."""
    lib().orc_set_num_threads(int(n))
    return int(lib().orc_get_max_threads())


def use_reference_kernel(on=True):
    """Route the restated loop nests through the reference's own coulomb_repulsion."""
    r = ref_lib()
    if on and r is None:
        raise RuntimeError("oracle/_ref/libpyquante2_ref.so is not available")
    ptr = C.cast(r.coulomb_repulsion, C.c_void_p) if on else None
    lib().orc_set_reference_kernel(ptr)


def ref_coulomb_repulsion(ra, rb, rc, rd, powers, alphas, norms, fn="coulomb_repulsion"):
    """Call the reference's coulomb_repulsion (or hrr) with pyquante2_coulomb_repulsion's
    argument shuffle (tho66.rs:237-259)."""
    r = ref_lib()
    args = []
    for k, ctr in enumerate((ra, rb, rc, rd)):
        args += [float(ctr[0]), float(ctr[1]), float(ctr[2]), float(norms[k]),
                 int(powers[3 * k]), int(powers[3 * k + 1]), int(powers[3 * k + 2]),
                 float(alphas[k])]
    return getattr(r, fn)(*args)


def coulomb_repulsion(ra, rb, rc, rd, powers, alphas, norms, boys=BOYS_REFERENCE):
    f = lambda v: np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_coulomb_repulsion(f(ra), f(rb), f(rc), f(rd),
                                       np.ascontiguousarray(powers, dtype=np.int32), f(alphas),
                                       f(norms), boys)


# ---------------------------------------------------------------------------------------
# data model (independent statement of basis.rs:108-212 and shell.rs:1-12)
# ---------------------------------------------------------------------------------------
BASIS_TABLES = {
    "STO-3G": {
        1: [([0], [3.42525091, 0.62391373, 0.16885540], [[0.15432897, 0.53532814, 0.44463454]])],
        8: [([0], [130.70932, 23.808861, 6.4436083], [[0.15432897, 0.53532814, 0.44463454]]),
            ([0, 1], [5.0331513, 1.1695961, 0.3803890],
             [[-0.09996723, 0.39951283, 0.70011547], [0.15591627, 0.60768372, 0.39195739]])],
    },
    "6-31G": {
        1: [([0], [18.7311370, 2.8253937, 0.6401217], [[0.03349460, 0.23472695, 0.81375733]]),
            ([0], [0.1612778], [[1.0]])],
        8: [([0], [5484.6717000, 825.2349500, 188.0469600, 52.9645000, 16.8975700, 5.7996353],
             [[0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.4701930, 0.3585209]]),
            ([0, 1], [15.5396160, 3.5999336, 1.0137618],
             [[-0.1107775, -0.1480263, 1.1307670], [0.0708743, 0.3397528, 0.7271586]]),
            ([0, 1], [0.2700058], [[1.0], [1.0]])],
    },
}
BASIS_TABLES["6-31G*"] = {
    1: BASIS_TABLES["6-31G"][1],
    8: BASIS_TABLES["6-31G"][8] + [([2], [0.8], [[1.0]])],
}


def ijk_list(m):
    out = np.zeros(((m + 1) * (m + 2) // 2, 3), dtype=np.int32)
    n = lib().orc_ijk_list(m, out.reshape(-1))
    assert n == out.shape[0]
    return out


class FlatBasis:
    """One entry per CGTO (= Cartesian function), flat arrays as the C ABI takes them."""

    def __init__(self, origins, powers, prim_offset, exps, coefs, norms):
        self.origins = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        self.powers = np.ascontiguousarray(powers, dtype=np.int32).reshape(-1, 3)
        self.prim_offset = np.ascontiguousarray(prim_offset, dtype=np.int32)
        self.exps = np.ascontiguousarray(exps, dtype=np.float64)
        self.coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        self.norms = np.ascontiguousarray(norms, dtype=np.float64)
        self.n = self.origins.shape[0]

    def args(self):
        return (self.n, self.origins.reshape(-1), self.powers.reshape(-1), self.prim_offset,
                self.exps, self.coefs, self.norms)


def make_basis(atomnos, coords, name):
    """Basis::new (basis.rs:182-211): atom -> shell -> angular momentum -> component."""
    table = BASIS_TABLES[name.upper().replace("(D)", "*")]
    origins, powers, off, exps, coefs, norms = [], [], [0], [], [], []
    L = lib()
    for z, xyz in zip(atomnos, np.asarray(coords, dtype=np.float64).reshape(-1, 3)):
        for ams, shell_exps, shell_coefs in table[int(z)]:
            for k, am in enumerate(ams):
                for pw in ijk_list(am):
                    origins.append(xyz)
                    powers.append(pw)
                    for e, c in zip(shell_exps, shell_coefs[k]):
                        exps.append(e)
                        coefs.append(c)
                        norms.append(L.orc_normalization(np.ascontiguousarray(pw), e))
                    off.append(len(exps))
    return FlatBasis(origins, powers, off, exps, coefs, norms)


def build_I(basis, boys=BOYS_REFERENCE):
    I = np.zeros((basis.n,) * 4)
    lib().orc_build_I(*basis.args(), I.reshape(-1), boys)
    return I


def jk_direct(basis, D, boys=BOYS_REFERENCE):
    J = np.zeros((basis.n, basis.n))
    K = np.zeros((basis.n, basis.n))
    lib().orc_jk_direct(*basis.args(), np.ascontiguousarray(D, dtype=np.float64).reshape(-1),
                        J.reshape(-1), K.reshape(-1), boys)
    return J, K


def jk_inmem(I, D):
    n = D.shape[0]
    J = np.zeros((n, n))
    K = np.zeros((n, n))
    lib().orc_jk_inmem(n, np.ascontiguousarray(I).reshape(-1),
                       np.ascontiguousarray(D, dtype=np.float64).reshape(-1), J.reshape(-1),
                       K.reshape(-1))
    return J, K


def one_electron(basis, which, atomnos=None, coords=None):
    """basis::S / T / V (basis.rs:234-338) with exact Boys (os86 path): which = 'S', 'T', 'V'"""
    M = np.zeros((basis.n, basis.n))
    z = np.ascontiguousarray(atomnos if atomnos is not None else [0], dtype=np.float64)
    xyz = np.ascontiguousarray(coords if coords is not None else [[0, 0, 0]], dtype=np.float64).reshape(-1)
    lib().orc_one_electron(*basis.args(), "STV".index(which), len(z) if which == "V" else 0, xyz, z,
                           M.reshape(-1))
    return M


def contracted_eri(basis, mu, nu, la, si, boys=BOYS_REFERENCE):
    return lib().orc_contracted_eri(*basis.args(), mu, nu, la, si, boys)


def eval_quartets(basis, quartets, boys=BOYS_REFERENCE, want_values=True):
    q = np.ascontiguousarray(quartets, dtype=np.int32).reshape(-1, 4)
    vals = np.zeros(q.shape[0]) if want_values else None
    ptr = vals.ctypes.data_as(C.c_void_p) if want_values else None
    total = lib().orc_eval_quartets(*basis.args(), q.reshape(-1), q.shape[0], ptr, boys)
    return vals if want_values else total


def quartet_list(Qb, Qk, same, tau):
    Qb = np.ascontiguousarray(Qb, dtype=np.float64)
    Qk = np.ascontiguousarray(Qk, dtype=np.float64)
    n = lib().orc_quartet_list(Qb, len(Qb), Qk, len(Qk), int(same), tau, None, 0)
    out = np.zeros((n, 2), dtype=np.int32)
    if n:
        lib().orc_quartet_list(Qb, len(Qb), Qk, len(Qk), int(same), tau,
                               out.ctypes.data_as(C.c_void_p), n)
    return out
