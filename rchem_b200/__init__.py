"""rchem_b200 -- B200-native two-electron-integral hot path of berquist/rchem.

Host-side mirror of the reference's public interface for this path (src/basis.rs), bound
over the C ABI of ``librchem_b200.so`` (include/rchem_eri.h) with ctypes:

    Basis.new(atomnos, coords, name)      <- basis::Basis::new        basis.rs:182-211
    S(basis), T(basis), V(basis, xyz, Z)  <- basis::S / T / V         basis.rs:234-338
    build_I(basis)                        <- basis::build_I           basis.rs:430-460
    JK_direct(J, K, basis, D)             <- basis::JK_direct         basis.rs:383-428
    JK_inmem(I, D)                        <- basis::JK_inmem          basis.rs:462-484
    get_ijk_list(m)                       <- shell::get_ijk_list      shell.rs:1-12
    coulomb_repulsion(...)                <- tho66::pyquante2::pyquante2_coulomb_repulsion
                                                                      tho66.rs:237-259

Same names, argument meaning and ordering as the reference; errors are Python exceptions
(``RchemError``) where the reference would panic/abort.  All compute happens in the CUDA
library: there is no CPU fallback, and importing this package fails loudly if the library
is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RCHEM_B200_LIB: development only (A/B timing of two builds of the same library)
LIB_PATH = os.environ.get("RCHEM_B200_LIB") or os.path.join(_HERE, "librchem_b200.so")

BOYS_REFERENCE, BOYS_EXACT = 0, 1
OPT_BOYS, OPT_SCHWARZ_TAU, OPT_DEVICE, OPT_PRIM_EPS, OPT_FAR_SCHED = 1, 2, 3, 4, 5
OPT_HEAVY_PASSES, OPT_SYMMETRIC_D_ONLY, OPT_LIGHT_KERNEL, OPT_NGPUS, OPT_FUSE_SP = 6, 7, 8, 9, 10
SHELL_SP = -1  # l reported for a fused s+p shell (4 functions: s, px, py, pz)


def shell_nfun(l):
    """functions of a shell whose rchem_basis_shells code is l"""
    l = int(l)
    return 4 if l == SHELL_SP else (l + 1) * (l + 2) // 2


class RchemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rchem error {code}: {msg}")
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C rchem_b200/csrc -j8` (there is no CPU fallback)")

_lib = C.CDLL(LIB_PATH)
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_vp = C.c_void_p


class Stats(C.Structure):
    _fields_ = [("shell_quartets", C.c_int64), ("shell_quartets_all", C.c_int64),
                ("prim_quartets", C.c_int64), ("integrals", C.c_int64),
                ("model_flops", C.c_double), ("kernel_ms", C.c_double),
                ("launches", C.c_int32), ("n_tasks", C.c_int32), ("setup_ms", C.c_double),
                ("fused_quartets", C.c_int64), ("prim_quartets_evaluated", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _sig(name, restype, argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_sig("rchem_last_error", C.c_char_p, [])
_sig("rchem_version", C.c_int, [])
_sig("rchem_device_count", C.c_int, [])
_sig("rchem_basis_new", C.c_int, [C.c_int, C.POINTER(C.c_uint64), _dp, C.c_char_p, C.POINTER(_vp)])
_sig("rchem_basis_create", C.c_int, [C.c_int, _dp, _ip, _ip, _dp, _dp, _dp, C.POINTER(_vp)])
_sig("rchem_basis_destroy", None, [_vp])
_sig("rchem_basis_nbf", C.c_int, [_vp])
_sig("rchem_basis_nprim", C.c_int, [_vp])
_sig("rchem_basis_nshells", C.c_int, [_vp])
_sig("rchem_basis_export", C.c_int, [_vp, _dp, _ip, _ip, _dp, _dp, _dp])
_sig("rchem_basis_shells", C.c_int, [_vp, _ip, _ip])
_sig("rchem_normalization", C.c_double, [_ip, C.c_double])
_sig("rchem_get_ijk_list", C.c_int, [C.c_int, _vp])
_sig("rchem_ijkl2intindex", C.c_int64, [C.c_int64] * 4)
_sig("rchem_set_option", C.c_int, [_vp, C.c_int, C.c_double])
_sig("rchem_get_option", C.c_double, [_vp, C.c_int])
_sig("rchem_set_stream", C.c_int, [_vp, _vp])
_sig("rchem_use_own_stream", C.c_int, [_vp])
_sig("rchem_build_I", C.c_int, [_vp, _dp])
_sig("rchem_jk_direct", C.c_int, [_vp, _dp, _dp, _dp])
_sig("rchem_jk_inmem", C.c_int, [C.c_int, _dp, _dp, _dp, _dp])
_sig("rchem_build_I_device", C.c_int, [_vp, _vp])
_sig("rchem_jk_direct_device", C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int])
_sig("rchem_jk_inmem_device", C.c_int, [C.c_int, _vp, _vp, _vp, _vp])
_sig("rchem_overlap", C.c_int, [_vp, _dp])
_sig("rchem_kinetic", C.c_int, [_vp, _dp])
_sig("rchem_nuclear", C.c_int, [_vp, C.c_int, _dp, C.POINTER(C.c_uint64), _dp])
_sig("rchem_schwarz", C.c_int64, [_vp, _vp, _vp, _vp, _vp])
_sig("rchem_quartet_list", C.c_int64, [_vp, _vp, C.c_int64])
_sig("rchem_get_stats", C.c_int, [_vp, C.POINTER(Stats)])
_sig("rchem_fp64_peak", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)])
_sig("rchem_coulomb_repulsion_batch", C.c_int, [C.c_int64, _dp, _dp, _ip, _dp, C.c_int, _dp])
_sig("coulomb_repulsion", C.c_double, ([C.c_double] * 4 + [C.c_int] * 3 + [C.c_double]) * 4)


def _check(rc):
    if rc < 0:
        raise RchemError(rc, (_lib.rchem_last_error() or b"").decode())
    return rc


def device_count():
    return _lib.rchem_device_count()


def fp64_peak(device=0, repeats=10):
    """Measured FP64-pipe peak (TFLOP/s): (best launch, average over `repeats` launches)."""
    best, avg = C.c_double(), C.c_double()
    _check(_lib.rchem_fp64_peak(device, repeats, C.byref(best), C.byref(avg)))
    return best.value, avg.value


def get_ijk_list(m):
    """shell::get_ijk_list (shell.rs:1-12)"""
    n = (m + 1) * (m + 2) // 2
    out = np.zeros((n, 3), dtype=np.int32)
    _check(_lib.rchem_get_ijk_list(m, out.ctypes.data_as(_vp)))
    return out


def normalization(powers, exponent):
    """PGTO::normalization (basis.rs:140-149)"""
    return _lib.rchem_normalization(np.ascontiguousarray(powers, dtype=np.int32), float(exponent))


def ijkl2intindex(i, j, k, l):
    """ijkl2intindex (cints.c:375-387), 64-bit"""
    return _lib.rchem_ijkl2intindex(i, j, k, l)


class Basis:
    """basis::Basis (basis.rs:175-212).  ``len(basis)`` is ``cgtos.len()``."""

    def __init__(self, handle):
        self._h = _vp(handle)
        self.nbf = _lib.rchem_basis_nbf(self._h)

    @classmethod
    def new(cls, atomnos, all_atomcoords, basis_set_name):
        """Basis::new(&atomnos, &all_atomcoords, basis_set_name); coordinates used raw."""
        z = np.ascontiguousarray(atomnos, dtype=np.uint64)
        xyz = np.ascontiguousarray(all_atomcoords, dtype=np.float64).reshape(-1, 3)
        if len(z) != xyz.shape[0]:
            raise ValueError("atomnos and all_atomcoords differ in length")
        h = _vp()
        _check(_lib.rchem_basis_new(len(z), z.ctypes.data_as(C.POINTER(C.c_uint64)),
                                    xyz.reshape(-1), basis_set_name.encode(), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_cgtos(cls, origins, powers, prim_offset, exps, coefs, norms):
        """One entry per CGTO (= Cartesian function): the fields of CGTO/PGTO, flattened."""
        origins = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1)
        n = origins.size // 3
        h = _vp()
        _check(_lib.rchem_basis_create(
            n, origins, np.ascontiguousarray(powers, dtype=np.int32).reshape(-1),
            np.ascontiguousarray(prim_offset, dtype=np.int32),
            np.ascontiguousarray(exps, dtype=np.float64),
            np.ascontiguousarray(coefs, dtype=np.float64),
            np.ascontiguousarray(norms, dtype=np.float64), C.byref(h)))
        return cls(h.value)

    def __len__(self):
        return self.nbf

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:  # module globals may already be gone at interpreter exit
            try:
                _lib.rchem_basis_destroy(h)
            except Exception:
                pass

    # --- data model access -----------------------------------------------------------
    def export(self):
        """(origins[n,3], powers[n,3], prim_offset[n+1], exps, coefs, norms)"""
        n, npr = self.nbf, _lib.rchem_basis_nprim(self._h)
        origins, powers = np.zeros(3 * n), np.zeros(3 * n, dtype=np.int32)
        off = np.zeros(n + 1, dtype=np.int32)
        exps, coefs, norms = np.zeros(npr), np.zeros(npr), np.zeros(npr)
        _check(_lib.rchem_basis_export(self._h, origins, powers, off, exps, coefs, norms))
        return origins.reshape(n, 3), powers.reshape(n, 3), off, exps, coefs, norms

    def shells(self):
        ns = _lib.rchem_basis_nshells(self._h)
        l, first = np.zeros(ns, dtype=np.int32), np.zeros(ns, dtype=np.int32)
        _check(_lib.rchem_basis_shells(self._h, l, first))
        return l, first

    # --- options ---------------------------------------------------------------------
    def set_boys(self, mode):
        _check(_lib.rchem_set_option(self._h, OPT_BOYS, float(mode)))

    def set_schwarz_tau(self, tau):
        _check(_lib.rchem_set_option(self._h, OPT_SCHWARZ_TAU, float(tau)))

    def set_prim_eps(self, eps):
        """Primitive-pair prefactor cutoff (default 1e-20; 0 keeps every primitive pair)."""
        _check(_lib.rchem_set_option(self._h, OPT_PRIM_EPS, float(eps)))

    def set_far_sched(self, on):
        """Far-field scheduling of the J/K kernels (default on; off = general code only)."""
        _check(_lib.rchem_set_option(self._h, OPT_FAR_SCHED, 1.0 if on else 0.0))

    def set_heavy_passes(self, passes):
        """J/K kernel choice (RCHEM_OPT_HEAVY_PASSES): bra pairs whose ket prefix fills the block
        kernel's threads `passes` times go to the block kernel (0 = every bra pair; a huge
        value = none, i.e. the warp-per-bra-pair / chunk kernels).  Same results."""
        _check(_lib.rchem_set_option(self._h, OPT_HEAVY_PASSES, float(passes)))

    def set_light_kernel(self, on):
        """Light bra pairs: warp-per-bra-pair kernel (default) or the chunk kernel."""
        _check(_lib.rchem_set_option(self._h, OPT_LIGHT_KERNEL, 1.0 if on else 0.0))

    def set_symmetric_only(self, on):
        """Reject an asymmetric D in JK_direct instead of paying a second build."""
        _check(_lib.rchem_set_option(self._h, OPT_SYMMETRIC_D_ONLY, 1.0 if on else 0.0))

    def set_fuse_sp(self, on):
        """RCHEM_OPT_FUSE_SP (before the first compute call): keep sp shells fused (default) or
        treat their s and p parts as separate shells.  Same integrals."""
        _check(_lib.rchem_set_option(self._h, OPT_FUSE_SP, 1.0 if on else 0.0))

    def set_gpus(self, n):
        """RCHEM_OPT_NGPUS: JK_direct drives n GPUs of this node from the one call (single
        process; devices device .. device+n-1; partial J/K summed over NVLink on the first)."""
        _check(_lib.rchem_set_option(self._h, OPT_NGPUS, float(n)))

    def set_device(self, ordinal):
        _check(_lib.rchem_set_option(self._h, OPT_DEVICE, float(ordinal)))

    def set_stream(self, cuda_stream_ptr):
        """Run on this cudaStream_t (0/None = the legacy default stream, which is what
        torch.cuda.current_stream().cuda_stream returns by default)."""
        _check(_lib.rchem_set_stream(self._h, _vp(cuda_stream_ptr or None)))

    def use_own_stream(self):
        _check(_lib.rchem_use_own_stream(self._h))

    def stats(self):
        s = Stats()
        _check(_lib.rchem_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    # --- screening -------------------------------------------------------------------
    def schwarz(self):
        """Shell pairs in kernel order: (shell_a, shell_b, batch, Q)."""
        n = _check(_lib.rchem_schwarz(self._h, None, None, None, None))
        a, b, bt = (np.zeros(n, dtype=np.int32) for _ in range(3))
        q = np.zeros(n)
        _check(_lib.rchem_schwarz(self._h, a.ctypes.data_as(_vp), b.ctypes.data_as(_vp),
                                  bt.ctypes.data_as(_vp), q.ctypes.data_as(_vp)))
        return a, b, bt, q

    def quartet_list(self):
        """Screened canonical quartet list as (pair_p, pair_q) rows (device-built)."""
        n = _check(_lib.rchem_quartet_list(self._h, None, 0))
        out = np.zeros((n, 2), dtype=np.int64)
        if n:
            _check(_lib.rchem_quartet_list(self._h, out.ctypes.data_as(_vp), n))
        return out

    # --- device-buffer entry points (raw pointers, e.g. torch tensor .data_ptr()) -------
    def jk_direct_device(self, D_ptr, JK_ptr, rank=0, nranks=1):
        """Asynchronous on the handle's stream: call set_stream() with the stream that produces
        D and consumes JK first (parallel.jk_direct_distributed does).  D must be symmetric."""
        _check(_lib.rchem_jk_direct_device(self._h, _vp(D_ptr), _vp(JK_ptr), rank, nranks))

    def build_I_device(self, I_ptr):
        _check(_lib.rchem_build_I_device(self._h, _vp(I_ptr)))


def S(basis_set):
    """basis::S(&basis_set): overlap matrix (basis.rs:234-251)."""
    out = np.zeros((basis_set.nbf, basis_set.nbf))
    _check(_lib.rchem_overlap(basis_set._h, out.reshape(-1)))
    return out


def T(basis_set):
    """basis::T(&basis_set): kinetic-energy matrix (basis.rs:273-290)."""
    out = np.zeros((basis_set.nbf, basis_set.nbf))
    _check(_lib.rchem_kinetic(basis_set._h, out.reshape(-1)))
    return out


def V(basis_set, atomcoords, atomnos):
    """basis::V(&basis_set, &atomcoords, &atomnos): nuclear-attraction matrix (basis.rs:316-338)."""
    xyz = np.ascontiguousarray(atomcoords, dtype=np.float64).reshape(-1, 3)
    z = np.ascontiguousarray(atomnos, dtype=np.uint64)
    out = np.zeros((basis_set.nbf, basis_set.nbf))
    _check(_lib.rchem_nuclear(basis_set._h, len(z), xyz.reshape(-1),
                              z.ctypes.data_as(C.POINTER(C.c_uint64)), out.reshape(-1)))
    return out


def build_I(basis_set):
    """basis::build_I(&basis_set) -> dense (N,N,N,N) array (basis.rs:430-460)."""
    n = basis_set.nbf
    out = np.zeros((n, n, n, n))
    _check(_lib.rchem_build_I(basis_set._h, out.reshape(-1)))
    return out


def JK_direct(J, K, basis_set, D):
    """basis::JK_direct(&mut J, &mut K, &basis_set, &D): J and K are overwritten in place
    (basis.rs:383-428).  Any D is accepted; one that is not symmetric costs a second build."""
    n = basis_set.nbf
    for name, m in (("J", J), ("K", K)):
        if m.shape != (n, n) or m.dtype != np.float64 or not m.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name} must be a C-contiguous float64 ({n},{n}) array")
    Dc = np.ascontiguousarray(D, dtype=np.float64)
    if Dc.shape != (n, n):
        raise ValueError(f"D must be ({n},{n})")
    _check(_lib.rchem_jk_direct(basis_set._h, Dc.reshape(-1), J.reshape(-1), K.reshape(-1)))


def JK_inmem(I, D):
    """basis::JK_inmem(&I, &D) -> (J, K) (basis.rs:462-484)."""
    n = D.shape[0]
    Ic = np.ascontiguousarray(I, dtype=np.float64)
    if Ic.shape != (n, n, n, n):
        raise ValueError("I must be (N,N,N,N)")
    J, K = np.zeros((n, n)), np.zeros((n, n))
    _check(_lib.rchem_jk_inmem(n, Ic.reshape(-1), np.ascontiguousarray(D, dtype=np.float64).reshape(-1),
                               J.reshape(-1), K.reshape(-1)))
    return J, K


def jk_inmem_device(n, I_ptr, D_ptr, JK_ptr, stream_ptr=None):
    _check(_lib.rchem_jk_inmem_device(n, _vp(I_ptr), _vp(D_ptr), _vp(JK_ptr), _vp(stream_ptr)))


def coulomb_repulsion(za, zb, zc, zd, ra, rb, rc, rd, norma, normb, normc, normd, c):
    """tho66::pyquante2::pyquante2_coulomb_repulsion (tho66.rs:237-259): same argument order."""
    v = _lib.coulomb_repulsion(
        ra[0], ra[1], ra[2], norma, c[0], c[1], c[2], za, rb[0], rb[1], rb[2], normb, c[3], c[4],
        c[5], zb, rc[0], rc[1], rc[2], normc, c[6], c[7], c[8], zc, rd[0], rd[1], rd[2], normd,
        c[9], c[10], c[11], zd)
    if v != v:
        raise RchemError(-1, (_lib.rchem_last_error() or b"").decode())
    return v


def coulomb_repulsion_batch(centres, norms, powers, alphas, boys=BOYS_REFERENCE):
    centres = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 12)
    n = centres.shape[0]
    out = np.zeros(n)
    _check(_lib.rchem_coulomb_repulsion_batch(
        n, centres.reshape(-1), np.ascontiguousarray(norms, dtype=np.float64).reshape(-1),
        np.ascontiguousarray(powers, dtype=np.int32).reshape(-1),
        np.ascontiguousarray(alphas, dtype=np.float64).reshape(-1), boys, out))
    return out
