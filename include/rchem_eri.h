/*
 * rchem_eri.h -- C ABI of the B200-native ERI hot path (librchem_b200.so).
 *
 * Drop-in boundary for the two-electron-integral path of berquist/rchem.  Each entry point
 * names the reference interface it replaces (paths relative to the reference checkout).
 * Plain pointers and sizes only; the caller owns every input/output buffer, the library
 * owns the opaque handle and the device memory behind it.
 *
 * Error behaviour: the reference aborts (C assert / exit(1), Rust unwrap panics); every
 * function here returns an int status instead (0 = ok, negative = error) and never aborts.
 * rchem_last_error() returns a message for the calling thread's last failure.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a) and
 * fails with RCHEM_ERR_NO_DEVICE / RCHEM_ERR_CUDA otherwise.
 *
 * Threading: a handle is not thread-safe (the reference is single-threaded); use one
 * handle per host thread.  One CUDA stream per handle.
 */
#ifndef RCHEM_ERI_H
#define RCHEM_ERI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCHEM_OK 0
#define RCHEM_ERR_INVALID_ARG (-1)
#define RCHEM_ERR_UNSUPPORTED_AM (-2)     /* angular momentum > 2 in a class kernel        */
#define RCHEM_ERR_UNSUPPORTED_LAYOUT (-3) /* CGTO list does not form Cartesian shells      */
#define RCHEM_ERR_CUDA (-4)
#define RCHEM_ERR_NO_DEVICE (-5)
#define RCHEM_ERR_ASYMMETRIC_D (-6)       /* JK needs a symmetric density (see rchem_jk_direct) */
#define RCHEM_ERR_TOO_LARGE (-7)          /* dense N^4 tensor does not fit device memory   */
#define RCHEM_ERR_UNKNOWN_BASIS (-8)

typedef struct rchem_basis rchem_basis; /* opaque: Basis (basis.rs:175-179) + device state */

const char* rchem_last_error(void);
int rchem_version(void);
/* number of visible CUDA devices, or a negative error */
int rchem_device_count(void);

/* ---------------- data model (src/basis.rs, src/shell.rs) ---------------------------- */

/* Basis::new(atomnos, all_atomcoords, basis_set_name)            basis.rs:182-211
 * coords are used raw (no unit conversion, like rchem.rs:29-35).  Basis data come from the
 * embedded tables (STO-3G, 6-31G, 6-31G* for H and O) instead of basis_set_exchange. */
int rchem_basis_new(int natoms, const uint64_t* atomnos, const double* coords,
                    const char* basis_set_name, rchem_basis** out);

/* Build a Basis from explicit CGTOs: one entry per Cartesian function, exactly the fields of
 * CGTO/PGTO (basis.rs:116-158): origins[3n], powers[3n], prim_offset[n+1], and per primitive
 * exponent, contraction coefficient and norm. */
int rchem_basis_create(int n_cgto, const double* origins, const int32_t* powers,
                       const int32_t* prim_offset, const double* exps, const double* coefs,
                       const double* norms, rchem_basis** out);
void rchem_basis_destroy(rchem_basis* b);

int rchem_basis_nbf(const rchem_basis* b);     /* basis_set.cgtos.len()                   */
int rchem_basis_nprim(const rchem_basis* b);   /* total primitives over all CGTOs         */
int rchem_basis_nshells(const rchem_basis* b); /* shells re-derived from the CGTO list    */
/* copies the flat CGTO arrays out (same layout as rchem_basis_create) */
int rchem_basis_export(const rchem_basis* b, double* origins, int32_t* powers,
                       int32_t* prim_offset, double* exps, double* coefs, double* norms);
/* per shell: angular momentum and index of its first function.  With RCHEM_OPT_FUSE_SP (default)
 * an s shell and a p shell on the same centre and exponents -- one "sp" electron shell of the
 * Basis Set Exchange data Basis::new reads (basis.rs:190-201) -- are ONE shell here, reported
 * with l = RCHEM_SHELL_SP: four functions s, px, py, pz. */
#define RCHEM_SHELL_SP (-1)
int rchem_basis_shells(const rchem_basis* b, int32_t* l, int32_t* first_function);

/* PGTO::normalization                                             basis.rs:140-149 */
double rchem_normalization(const int32_t powers[3], double exponent);
/* shell::get_ijk_list(m): writes 3*(m+1)(m+2)/2 ints, returns the count   shell.rs:1-12 */
int rchem_get_ijk_list(int m, int32_t* out);
/* ijkl2intindex, 64-bit (the reference's int overflows past N=361) cints.c:375-387 */
int64_t rchem_ijkl2intindex(int64_t i, int64_t j, int64_t k, int64_t l);

/* ---------------- options ------------------------------------------------------------ */
#define RCHEM_OPT_BOYS 1        /* 0 = reference Fgamma (default; matches libpyquante2 to 1e-12),
                                   1 = exact Boys (~1e-15; differs from the reference by <=2e-8) */
#define RCHEM_OPT_SCHWARZ_TAU 2 /* drop shell quartets with Q_ab*Q_cd < tau; default 0 = none
                                   (the reference does not screen)                  */
#define RCHEM_OPT_DEVICE 3      /* CUDA device ordinal, before the first compute call          */
#define RCHEM_OPT_PRIM_EPS 4    /* drop primitive pairs with |c_a c_b N_a N_b exp(..)/zeta| below
                                   this (default 1e-20: < 1e-15 per integral; 0 keeps all);
                                   before the first compute call                               */
#define RCHEM_OPT_FAR_SCHED 5   /* 1 (default): J/K kernels prove shell quartets far-field from the
                                   pairs' bounding spheres and evaluate those by the point-
                                   multipole form; 0: every quartet through the general code
                                   (same results to ~1e-15; a tuning / cross-check knob)        */
#define RCHEM_OPT_HEAVY_PASSES 6 /* J/K kernel choice: a bra pair whose surviving ket prefix fills
                                   the block kernel's threads at least this many times (default 2)
                                   goes to the block-per-bra-pair kernel, the others to the
                                   warp-per-bra-pair kernel; 0 sends every bra pair with >= 1 ket to
                                   the block kernel.  Same results either way (a tuning knob, and
                                   how the tests force each kernel).                            */
#define RCHEM_OPT_LIGHT_KERNEL 8 /* 1 (default): bra pairs below the heavy threshold use the
                                   warp-per-bra-pair kernel; 0: the warp-per-32-kets chunk kernel
                                   (the fallback of classes without a block kernel)            */
#define RCHEM_OPT_SYMMETRIC_D_ONLY 7 /* 1: rchem_jk_direct rejects an asymmetric D with
                                   RCHEM_ERR_ASYMMETRIC_D instead of paying a second build
                                   (default 0: any D is accepted, like basis.rs:383-428)       */
#define RCHEM_OPT_FUSE_SP 10     /* 1 (default): bases made of s and sp shells only (STO-3G, 6-31G)
                                   keep each sp shell FUSED -- one shell quartet evaluates every
                                   s|p-part combination on one set of primitive quartets (shared
                                   Boys values and recurrences); 0: s and p parts as separate
                                   shells.  Same integrals.  Before the first compute call.    */
#define RCHEM_OPT_NGPUS 9        /* n > 1: rchem_jk_direct drives the n devices device .. device+n-1
                                   of this node from the ONE call (single process): H2D of D
                                   once, peer copies of D over NVLink, every device builds its
                                   block-interleaved share of the quartets, device `device` sums
                                   the partial [J|K] out of its peers' memory (peer-to-peer
                                   loads), one D2H.  Default 1.  An asymmetric D and every other
                                   entry point use device `device` alone.                      */
int rchem_set_option(rchem_basis* b, int key, double value);
double rchem_get_option(const rchem_basis* b, int key);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream).  The value is used as
 * is: NULL is the CUDA legacy default stream.  rchem_use_own_stream() goes back to the
 * handle's private non-blocking stream (the initial state). */
int rchem_set_stream(rchem_basis* b, void* cuda_stream);
int rchem_use_own_stream(rchem_basis* b);

/* ---------------- the hot path, HOST buffers (copies are inside the call) ------------- */

/* build_I(&basis) -> dense row-major I[mu][nu][la][si], N^4 doubles       basis.rs:430-460 */
int rchem_build_I(rchem_basis* b, double* I_host);
/* JK_direct(&mut J, &mut K, &basis, &D): J,K (N x N) are overwritten       basis.rs:383-428
 * Any D is accepted, like the reference.  The kernels exploit D = D^T (what the reference's
 * caller passes: D = C C^T, rchem.rs:101-104); a D that is not symmetric (checked on the
 * device, relative 1e-14) is split into its symmetric and antisymmetric parts and costs a
 * second build: J(D) = J(S), K(D) = K(S) + K(A). */
int rchem_jk_direct(rchem_basis* b, const double* D_host, double* J_host, double* K_host);
/* JK_inmem(&I, &D) -> (J, K)                                               basis.rs:462-484 */
int rchem_jk_inmem(int n, const double* I_host, const double* D_host, double* J_host,
                   double* K_host);

/* ---------------- the hot path, DEVICE buffers (no copies; asynchronous on the stream) - */
int rchem_build_I_device(rchem_basis* b, double* I_dev);
/* This rank's additive share of J and K: JK_dev holds J then K (2*N*N doubles).  With
 * nranks > 1 the caller sums JK_dev over ranks (one allreduce); rank 0 of 1 gives J, K.
 * D_dev MUST be symmetric here (no check, no copy: this is the asynchronous inner call;
 * rchem_jk_direct is the entry point that accepts any D).  The work is enqueued on the handle's
 * stream (rchem_set_stream): bind it to the stream that produces D_dev / consumes JK_dev. */
int rchem_jk_direct_device(rchem_basis* b, const double* D_dev, double* JK_dev, int rank,
                           int nranks);
int rchem_jk_inmem_device(int n, const double* I_dev, const double* D_dev, double* JK_dev,
                          void* cuda_stream);

/* ---------------- one-electron matrices: the step before the hot path ------------------
 * basis::S(&basis), basis::T(&basis), basis::V(&basis, &atomcoords, &atomnos)
 * (basis.rs:234-338): N x N, symmetric, host buffers.  Evaluated on the GPU with the exact
 * Boys function, like the reference's os86 path (os86.rs:629, 672, 722). */
int rchem_overlap(rchem_basis* b, double* S_host);
int rchem_kinetic(rchem_basis* b, double* T_host);
int rchem_nuclear(rchem_basis* b, int natoms, const double* atomcoords, const uint64_t* atomnos,
                  double* V_host);

/* ---------------- screening --------------------------------------------------------- */
/* Shell pairs in kernel order (batch by batch, Schwarz-descending inside a batch).
 * Returns the number of pairs; arrays may be NULL to query the count.
 *   shell_a/shell_b : shell indices (l_a >= l_b), batch : batch id, Q : sqrt(max|(ab|ab)|) */
int64_t rchem_schwarz(rchem_basis* b, int32_t* shell_a, int32_t* shell_b, int32_t* batch,
                      double* Q);
/* Canonical screened quartet list for the current tau, materialised on the device and copied
 * out as (pair_p, pair_q) indices into the rchem_schwarz arrays.  Returns the count; pass
 * pq = NULL to query.  For tests and small systems only. */
int64_t rchem_quartet_list(rchem_basis* b, int64_t* pq, int64_t capacity);

/* ---------------- statistics of the last build_I / jk_direct call -------------------- */
typedef struct {
  int64_t shell_quartets;     /* computed (after screening), this rank; counted in SEGMENTED
                                 (s/p/d) shell quartets: a quartet of fused sp shells covers
                                 every s|p-part combination of its four shells              */
  int64_t shell_quartets_all; /* canonical quartets before screening                       */
  int64_t prim_quartets;      /* primitive quartets of those segmented quartets, this rank  */
  int64_t integrals;          /* Cartesian (ab|cd) values produced, this rank              */
  double model_flops;         /* SURVEY 8(d) flop model summed over computed quartets      */
  double kernel_ms;           /* device time of the ERI kernels (CUDA events on the stream) */
  int32_t launches;           /* kernels launched (ERI + finalize)                         */
  int32_t n_tasks;            /* batch pairs                                               */
  double setup_ms;            /* host wall time spent so far on one-off set-up of this handle:
                                 pair batches, Schwarz bounds, Boys tables, task tables      */
  int64_t fused_quartets;     /* shell quartets as the kernels see them (fused sp shells = 1) */
  int64_t prim_quartets_evaluated; /* primitive quartets actually evaluated (shared by the
                                 variants of a fused quartet)                                */
} rchem_stats;
int rchem_get_stats(const rchem_basis* b, rchem_stats* out);

/* Measured FP64-pipe peak of `device` (DFMA microbenchmark, TFLOP/s): best of `repeats`
 * launches and the average over them.  The denominator of the FP64 roofline fraction. */
int rchem_fp64_peak(int device, int repeats, double* tflops_best, double* tflops_sustained);

/* ---------------- tier 1: symbol-compatible primitive integral ----------------------- */
/* coulomb_repulsion with the exact libpyquante2 signature (cints.h:23-30), evaluated on the
 * GPU (one launch per call: use the batch form for more than a handful).  Returns NaN on
 * error. */
double coulomb_repulsion(double xa, double ya, double za, double norma, int la, int ma, int na,
                         double alphaa, double xb, double yb, double zb, double normb, int lb,
                         int mb, int nb, double alphab, double xc, double yc, double zc,
                         double normc, int lc, int mc, int nc, double alphac, double xd,
                         double yd, double zd, double normd, int ld, int md, int nd,
                         double alphad);
/* n primitive integrals; per integral: centres[12], norms[4], powers[12], alphas[4] */
int rchem_coulomb_repulsion_batch(int64_t n, const double* centres, const double* norms,
                                  const int32_t* powers, const double* alphas, int boys,
                                  double* out);

#ifdef __cplusplus
}
#endif
#endif /* RCHEM_ERI_H */
