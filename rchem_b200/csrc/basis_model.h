// basis_model.h -- host-side mirror of the reference data model (src/basis.rs, src/shell.rs).
//
//   PGTO  {origin, powers, exponent, norm}            basis.rs:116-150
//   CGTO  {origin, powers, coefs, primitives}         basis.rs:152-173
//   Basis {name, cgtos}; Basis::new                   basis.rs:175-212
//   get_ijk_list                                      shell.rs:1-12
//
// One CGTO per Cartesian basis FUNCTION, in the reference's order atom -> BSE shell ->
// angular momentum -> get_ijk_list component (basis.rs:186-203).  The reference has no shell
// objects; group_shells() re-derives them for the kernels.
//
// The reference pulls basis sets from the Python package basis_set_exchange at run time
// (basis.rs:87-106); that package and the network are not available, so the tables for the
// configurations of BASELINE.json (H and O; STO-3G, 6-31G, 6-31G*) are embedded.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace rchem {

struct PGTO {
  double origin[3];
  int powers[3];
  double exponent;
  double norm;
};

struct CGTO {
  double origin[3];
  int powers[3];
  std::vector<double> coefs;
  std::vector<PGTO> primitives;
};

struct Basis {
  std::string name;
  std::vector<CGTO> cgtos;
};

// shell::get_ijk_list (shell.rs:1-12)
inline std::vector<std::array<int, 3>> get_ijk_list(int m) {
  std::vector<std::array<int, 3>> out;
  for (int a = 1; a < m + 2; ++a)
    for (int b = 1; b < a + 1; ++b) out.push_back({m + 1 - a, a - b, b - 1});
  return out;
}

// basis.rs:108-114
inline long fact2(long n) { return n <= 0 ? 1 : n * fact2(n - 2); }

// PGTO::normalization (basis.rs:140-149)
inline double pgto_normalization(const int powers[3], double exponent) {
  const double order = powers[0] + powers[1] + powers[2];
  const double numer = std::pow(2.0, 2.0 * order + 1.5) * std::pow(exponent, order + 1.5);
  const double denom =
      (double)(fact2(2 * powers[0] - 1) * fact2(2 * powers[1] - 1) * fact2(2 * powers[2] - 1)) *
      std::pow(M_PI, 1.5);
  return std::pow(numer / denom, 0.5);
}

// ---- embedded basis-set data, in Basis Set Exchange "electron_shells" layout ------------
struct BseShell {
  std::vector<int> angular_momentum;            // e.g. {0,1} for an sp shell
  std::vector<double> exponents;
  std::vector<std::vector<double>> coefficients;  // one row per angular momentum
};

// Returns nullptr when the (basis, element) pair is not embedded.
const std::vector<BseShell>* embedded_basis(const std::string& name, unsigned atomno);

// Basis::new (basis.rs:182-211).  Returns false + message for unknown basis/element.
bool basis_new(const std::vector<uint64_t>& atomnos, const double* coords,
               const std::string& basis_set_name, Basis* out, std::string* err);

// ---- shells re-derived from the CGTO list ----------------------------------------------
struct Shell {
  int l;       // shell TYPE (eri_core.h): angular momentum 0..4, or 3 = kTypeSP for a fused s+p
               // shell when ShellSet::fused (class kernels only go up to d, so 3 is never f there)
  int bf0;     // index of the first Cartesian function of the shell in Basis::cgtos
  double ctr[3];
  std::vector<double> exps;
  std::vector<double> cn;  // coefs[i] * norm of the FIRST component (L,0,0) of primitive i
  std::vector<double> cn2; // fused sp shell: the same for its p part (cn is the s part)
};

struct ShellSet {
  std::vector<Shell> shells;
  // compscale[l][k] = norm(component k)/norm(component 0); identical for every shell of
  // that l (checked), e.g. d: {1, sqrt3, sqrt3, 1, sqrt3, 1}
  double compscale[5][15];
  int lmax = 0;
  bool fused = false;  // s+p shells sharing centre and exponents were fused into type-3 shells
};

// Groups consecutive CGTOs into shells.  Fails (false + message) when the functions do not
// form complete Cartesian shells in get_ijk_list order with shared exponents/coefficients,
// or when the caller-supplied norms are not "shell norm x component factor".
// fuse_sp: additionally fuse every (s shell, p shell) pair of consecutive functions that share
// centre and exponents (the sp shells of STO-3G / 6-31G) into ONE shell of type kTypeSP -- done
// only when the basis consists of s and such sp shells alone (every p shell fuses, no d shell),
// because the class kernels exist for the type sets {s,p,d} and {s,sp}.
bool group_shells(const Basis& basis, ShellSet* out, std::string* err, bool fuse_sp = false);

}  // namespace rchem
