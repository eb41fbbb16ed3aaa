//! rchem-b200-sys -- Rust side of the drop-in boundary (`include/rchem_eri.h`).
//!
//! **UNVERIFIED SOURCE**: no Rust toolchain exists in the image this was written in
//!.  Every symbol bound here is exercised through the same C ABI by
//! `tests/abi_consumer.c` (gcc) and by the ctypes mirror `rchem_b200/__init__.py`.
//!
//! Two tiers, as in INTEGRATION.md:
//!
//! * tier 1 -- [`pyquante2`]: `coulomb_repulsion` with libpyquante2's exact signature
//!   (cints.h:23-30) and the safe wrapper of `src/integrals/tho66.rs:237-259`, now resolved
//!   from `librchem_b200.so` (evaluated on the GPU);
//! * tier 2 -- [`BasisHandle`] and the three builders the reference's `src/basis.rs` exposes:
//!   `JK_direct` (basis.rs:383-428), `build_I` (430-460), `JK_inmem` (462-484), plus the
//!   set-up matrices `S`/`T`/`V` (234-338).
#![allow(non_snake_case)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_double, c_int, c_void};

/// Opaque `rchem_basis` (the reference's `Basis` + device state).
#[repr(C)]
pub struct rchem_basis {
    _private: [u8; 0],
}

/// `rchem_stats` of include/rchem_eri.h (same field order).
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct rchem_stats {
    pub shell_quartets: i64,
    pub shell_quartets_all: i64,
    pub prim_quartets: i64,
    pub integrals: i64,
    pub model_flops: c_double,
    pub kernel_ms: c_double,
    pub launches: i32,
    pub n_tasks: i32,
    pub setup_ms: c_double,
    pub fused_quartets: i64,
    pub prim_quartets_evaluated: i64,
}

pub const RCHEM_OK: c_int = 0;
pub const RCHEM_OPT_BOYS: c_int = 1;
pub const RCHEM_OPT_SCHWARZ_TAU: c_int = 2;
pub const RCHEM_OPT_DEVICE: c_int = 3;
pub const RCHEM_OPT_PRIM_EPS: c_int = 4;
pub const RCHEM_OPT_FAR_SCHED: c_int = 5;
pub const RCHEM_OPT_HEAVY_PASSES: c_int = 6;
pub const RCHEM_OPT_SYMMETRIC_D_ONLY: c_int = 7;
pub const RCHEM_OPT_LIGHT_KERNEL: c_int = 8;
pub const RCHEM_OPT_NGPUS: c_int = 9;
pub const RCHEM_OPT_FUSE_SP: c_int = 10;
/// `l` reported by `rchem_basis_shells` for a fused s+p shell (s, px, py, pz).
pub const RCHEM_SHELL_SP: i32 = -1;

extern "C" {
    pub fn rchem_last_error() -> *const c_char;
    pub fn rchem_version() -> c_int;
    pub fn rchem_device_count() -> c_int;
    pub fn rchem_basis_new(natoms: c_int, atomnos: *const u64, coords: *const c_double,
                           basis_set_name: *const c_char, out: *mut *mut rchem_basis) -> c_int;
    pub fn rchem_basis_create(n_cgto: c_int, origins: *const c_double, powers: *const i32,
                              prim_offset: *const i32, exps: *const c_double,
                              coefs: *const c_double, norms: *const c_double,
                              out: *mut *mut rchem_basis) -> c_int;
    pub fn rchem_basis_destroy(b: *mut rchem_basis);
    pub fn rchem_basis_nbf(b: *const rchem_basis) -> c_int;
    pub fn rchem_set_option(b: *mut rchem_basis, key: c_int, value: c_double) -> c_int;
    pub fn rchem_get_option(b: *const rchem_basis, key: c_int) -> c_double;
    pub fn rchem_set_stream(b: *mut rchem_basis, cuda_stream: *mut c_void) -> c_int;
    pub fn rchem_build_I(b: *mut rchem_basis, i_host: *mut c_double) -> c_int;
    pub fn rchem_jk_direct(b: *mut rchem_basis, d_host: *const c_double, j_host: *mut c_double,
                           k_host: *mut c_double) -> c_int;
    pub fn rchem_jk_inmem(n: c_int, i_host: *const c_double, d_host: *const c_double,
                          j_host: *mut c_double, k_host: *mut c_double) -> c_int;
    pub fn rchem_overlap(b: *mut rchem_basis, s_host: *mut c_double) -> c_int;
    pub fn rchem_kinetic(b: *mut rchem_basis, t_host: *mut c_double) -> c_int;
    pub fn rchem_nuclear(b: *mut rchem_basis, natoms: c_int, atomcoords: *const c_double,
                         atomnos: *const u64, v_host: *mut c_double) -> c_int;
    pub fn rchem_get_stats(b: *const rchem_basis, out: *mut rchem_stats) -> c_int;
    pub fn rchem_ijkl2intindex(i: i64, j: i64, k: i64, l: i64) -> i64;
}

/// Error of a library call: the negative status and `rchem_last_error()`.
#[derive(Debug)]
pub struct Error {
    pub code: c_int,
    pub message: String,
}

fn check(code: c_int) -> Result<(), Error> {
    if code == RCHEM_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(rchem_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code, message })
}

/// Tier 1: drop-in for `src/integrals/tho66.rs:225-260` (`mod pyquante2`).  The reference
/// `include!`s bindgen output for cints.h; the one prototype it uses is declared here.
pub mod pyquante2 {
    use std::os::raw::{c_double, c_int};
    extern "C" {
        // cints.h:23-30, exported by librchem_b200.so with the identical signature
        pub fn coulomb_repulsion(
            xa: c_double, ya: c_double, za: c_double, norma: c_double, la: c_int, ma: c_int,
            na: c_int, alphaa: c_double, xb: c_double, yb: c_double, zb: c_double,
            normb: c_double, lb: c_int, mb: c_int, nb: c_int, alphab: c_double, xc: c_double,
            yc: c_double, zc: c_double, normc: c_double, lc: c_int, mc: c_int, nc: c_int,
            alphac: c_double, xd: c_double, yd: c_double, zd: c_double, normd: c_double,
            ld: c_int, md: c_int, nd: c_int, alphad: c_double,
        ) -> c_double;
    }

    /// Same argument meaning and order as the reference wrapper (tho66.rs:237-259): exponents,
    /// centres, norms, then the twelve Cartesian powers.
    pub fn pyquante2_coulomb_repulsion(
        za: f64, zb: f64, zc: f64, zd: f64, ra: &[f64; 3], rb: &[f64; 3], rc: &[f64; 3],
        rd: &[f64; 3], norma: f64, normb: f64, normc: f64, normd: f64, c: &[i32; 12],
    ) -> f64 {
        unsafe {
            coulomb_repulsion(
                ra[0], ra[1], ra[2], norma, c[0], c[1], c[2], za,
                rb[0], rb[1], rb[2], normb, c[3], c[4], c[5], zb,
                rc[0], rc[1], rc[2], normc, c[6], c[7], c[8], zc,
                rd[0], rd[1], rd[2], normd, c[9], c[10], c[11], zd,
            )
        }
    }
}

/// The flat CGTO arrays `rchem_basis_create` takes: one entry per Cartesian function, exactly
/// the fields of `CGTO`/`PGTO` (basis.rs:116-158).  `basis.rs` fills this from `Basis::cgtos`
/// (see rust/patches/basis_rs.patch).
#[derive(Default)]
pub struct FlatCgtos {
    pub origins: Vec<f64>,
    pub powers: Vec<i32>,
    pub prim_offset: Vec<i32>,
    pub exps: Vec<f64>,
    pub coefs: Vec<f64>,
    pub norms: Vec<f64>,
}

impl FlatCgtos {
    pub fn new() -> Self {
        FlatCgtos { prim_offset: vec![0], ..Default::default() }
    }
    /// Append one CGTO: origin, Cartesian powers, and per primitive (exponent, coef, norm).
    pub fn push(&mut self, origin: &[f64; 3], powers: &[usize; 3], prims: &[(f64, f64, f64)]) {
        self.origins.extend_from_slice(origin);
        self.powers.extend(powers.iter().map(|&p| p as i32));
        for &(e, c, n) in prims {
            self.exps.push(e);
            self.coefs.push(c);
            self.norms.push(n);
        }
        self.prim_offset.push(self.exps.len() as i32);
    }
    pub fn len(&self) -> usize {
        self.prim_offset.len() - 1
    }
}

/// Owning wrapper of the opaque handle.  Not `Sync`: one handle per host thread (the reference
/// is single-threaded).
pub struct BasisHandle {
    raw: *mut rchem_basis,
    n: usize,
}

impl BasisHandle {
    pub fn from_cgtos(f: &FlatCgtos) -> Result<Self, Error> {
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            rchem_basis_create(f.len() as c_int, f.origins.as_ptr(), f.powers.as_ptr(),
                               f.prim_offset.as_ptr(), f.exps.as_ptr(), f.coefs.as_ptr(),
                               f.norms.as_ptr(), &mut raw)
        })?;
        Ok(BasisHandle { raw, n: f.len() })
    }
    /// `Basis::new(atomnos, all_atomcoords, basis_set_name)` on the embedded tables.
    pub fn new(atomnos: &[u64], coords: &[[f64; 3]], name: &str) -> Result<Self, Error> {
        let cname = std::ffi::CString::new(name).unwrap();
        let mut raw = std::ptr::null_mut();
        check(unsafe {
            rchem_basis_new(atomnos.len() as c_int, atomnos.as_ptr(),
                            coords.as_ptr() as *const c_double, cname.as_ptr(), &mut raw)
        })?;
        let n = unsafe { rchem_basis_nbf(raw) } as usize;
        Ok(BasisHandle { raw, n })
    }
    pub fn nbf(&self) -> usize {
        self.n
    }
    pub fn set_option(&mut self, key: c_int, value: f64) -> Result<(), Error> {
        check(unsafe { rchem_set_option(self.raw, key, value) })
    }
    /// Drive `n` GPUs of this node from the one `JK_direct` call (RCHEM_OPT_NGPUS).
    pub fn set_gpus(&mut self, n: usize) -> Result<(), Error> {
        self.set_option(RCHEM_OPT_NGPUS, n as f64)
    }
    /// `JK_direct(&mut J, &mut K, &basis, &D)`: row-major N x N slices, J and K overwritten.
    pub fn jk_direct(&mut self, d: &[f64], j: &mut [f64], k: &mut [f64]) -> Result<(), Error> {
        let nn = self.n * self.n;
        assert!(d.len() == nn && j.len() == nn && k.len() == nn);
        check(unsafe { rchem_jk_direct(self.raw, d.as_ptr(), j.as_mut_ptr(), k.as_mut_ptr()) })
    }
    /// `build_I(&basis)`: row-major N^4 slice.
    pub fn build_i(&mut self, i: &mut [f64]) -> Result<(), Error> {
        assert!(i.len() == self.n.pow(4));
        check(unsafe { rchem_build_I(self.raw, i.as_mut_ptr()) })
    }
    pub fn overlap(&mut self, s: &mut [f64]) -> Result<(), Error> {
        check(unsafe { rchem_overlap(self.raw, s.as_mut_ptr()) })
    }
    pub fn kinetic(&mut self, t: &mut [f64]) -> Result<(), Error> {
        check(unsafe { rchem_kinetic(self.raw, t.as_mut_ptr()) })
    }
    pub fn nuclear(&mut self, coords: &[[f64; 3]], atomnos: &[u64], v: &mut [f64]) -> Result<(), Error> {
        check(unsafe {
            rchem_nuclear(self.raw, atomnos.len() as c_int, coords.as_ptr() as *const c_double,
                          atomnos.as_ptr(), v.as_mut_ptr())
        })
    }
    pub fn stats(&self) -> rchem_stats {
        let mut s = rchem_stats::default();
        unsafe { rchem_get_stats(self.raw, &mut s) };
        s
    }
}

impl Drop for BasisHandle {
    fn drop(&mut self) {
        unsafe { rchem_basis_destroy(self.raw) }
    }
}

/// `JK_inmem(&I, &D) -> (J, K)` on flat row-major slices (basis.rs:462-484).
pub fn jk_inmem(n: usize, i: &[f64], d: &[f64]) -> Result<(Vec<f64>, Vec<f64>), Error> {
    assert!(i.len() == n.pow(4) && d.len() == n * n);
    let (mut j, mut k) = (vec![0.0; n * n], vec![0.0; n * n]);
    check(unsafe { rchem_jk_inmem(n as c_int, i.as_ptr(), d.as_ptr(), j.as_mut_ptr(), k.as_mut_ptr()) })?;
    Ok((j, k))
}

/// ndarray front end with the reference's own signatures (`Array<f64, Ix2>` / `Ix4`, row-major,
/// which is ndarray's default layout and what the C ABI reads and writes).
#[cfg(feature = "ndarray")]
pub mod nd {
    use super::*;
    use ndarray::{Array, Ix2, Ix4};

    pub fn JK_direct(J: &mut Array<f64, Ix2>, K: &mut Array<f64, Ix2>, basis: &mut BasisHandle,
                     D: &Array<f64, Ix2>) -> Result<(), Error> {
        let d = D.as_standard_layout();
        basis.jk_direct(d.as_slice().unwrap(), J.as_slice_mut().unwrap(), K.as_slice_mut().unwrap())
    }
    pub fn build_I(basis: &mut BasisHandle) -> Result<Array<f64, Ix4>, Error> {
        let n = basis.nbf();
        let mut I = Array::zeros((n, n, n, n));
        basis.build_i(I.as_slice_mut().unwrap())?;
        Ok(I)
    }
    pub fn JK_inmem(I: &Array<f64, Ix4>, D: &Array<f64, Ix2>) -> Result<(Array<f64, Ix2>, Array<f64, Ix2>), Error> {
        let n = D.shape()[0];
        let (j, k) = jk_inmem(n, I.as_standard_layout().as_slice().unwrap(),
                              D.as_standard_layout().as_slice().unwrap())?;
        Ok((Array::from_shape_vec((n, n), j).unwrap(), Array::from_shape_vec((n, n), k).unwrap()))
    }
}

#[cfg(test)]
mod tests {
    use super::pyquante2::pyquante2_coulomb_repulsion;

    /// The reference's own golden value (tho66.rs:288-321), through the GPU symbol.
    #[test]
    fn test_coulomb_repulsion() {
        let c = [0i32; 12];
        let v = pyquante2_coulomb_repulsion(1.1, 1.2, 1.3, 1.4, &[1.0, 0.0, 1.0], &[0.0, 1.0, 2.0],
                                            &[0.0, 0.0, 3.0], &[0.0, 0.0, 4.0], 1.0, 1.0, 1.0, 1.0, &c);
        assert!((v - 0.08608517834596989).abs() < 1e-12);
    }
}
