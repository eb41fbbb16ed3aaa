// hostcheck.cpp -- TEST-ONLY host build of the per-quartet arithmetic (eri_core.h + the
// generated recurrences), so the recurrences, the Boys restatement, the prefactors and the
// component ordering can be checked against the oracle in the CPU-only container.
// It is compiled into tests/_build/libhostcheck.so by tests/conftest.py and is never part of,
// nor loaded by, the product library (which has no CPU path).
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../rchem_b200/csrc/basis_model.h"
#include "../rchem_b200/csrc/eri_core.h"
#include "../rchem_b200/csrc/pair_build.h"
#include "../rchem_b200/csrc/boys_delta_build.h"
#include "../rchem_b200/csrc/gen/eri_class_list.h"

namespace rchem {
#include "../rchem_b200/csrc/gen/eri_class_0000.inc"
#include "../rchem_b200/csrc/gen/eri_class_1000.inc"
#include "../rchem_b200/csrc/gen/eri_class_1010.inc"
#include "../rchem_b200/csrc/gen/eri_class_1100.inc"
#include "../rchem_b200/csrc/gen/eri_class_1110.inc"
#include "../rchem_b200/csrc/gen/eri_class_1111.inc"
#include "../rchem_b200/csrc/gen/eri_class_2000.inc"
#include "../rchem_b200/csrc/gen/eri_class_2010.inc"
#include "../rchem_b200/csrc/gen/eri_class_2011.inc"
#include "../rchem_b200/csrc/gen/eri_class_2020.inc"
#include "../rchem_b200/csrc/gen/eri_class_2100.inc"
#include "../rchem_b200/csrc/gen/eri_class_2110.inc"
#include "../rchem_b200/csrc/gen/eri_class_2111.inc"
#include "../rchem_b200/csrc/gen/eri_class_2120.inc"
#include "../rchem_b200/csrc/gen/eri_class_2121.inc"
#include "../rchem_b200/csrc/gen/eri_class_2200.inc"
#include "../rchem_b200/csrc/gen/eri_class_2210.inc"
#include "../rchem_b200/csrc/gen/eri_class_2211.inc"
#include "../rchem_b200/csrc/gen/eri_class_2220.inc"
#include "../rchem_b200/csrc/gen/eri_class_2221.inc"
#include "../rchem_b200/csrc/gen/eri_class_2222.inc"
#include "../rchem_b200/csrc/gen/eri_class_3000.inc"
#include "../rchem_b200/csrc/gen/eri_class_3030.inc"
#include "../rchem_b200/csrc/gen/eri_class_3300.inc"
#include "../rchem_b200/csrc/gen/eri_class_3330.inc"
#include "../rchem_b200/csrc/gen/eri_class_3333.inc"

// BOYS == kBoysFarForm (test-only value): every primitive quartet through the far-field form
// primitive_quartet_far, whatever its x; *min_x reports the smallest Boys argument met and
// *proved_far what the block kernel's bounding-sphere test (eri_kernel.cuh) says.
constexpr int kBoysFarForm = 4;
static int g_prims_used = 0;
static int g_fuse_sp = 0;  // hostcheck_set_fuse: group sp shells as ONE fused shell (type 3)
static double g_min_x = 0.0;
static int g_proved_far = 0;
template <class C, int BOYS>
void shell_quartet(const ShellSet& ss, const Shell& A, const Shell& B, const Shell& Cc,
                   const Shell& D, const BoysTabs& tabs, double* out) {
  std::vector<PrimPair> bra, ket;
  build_prim_pairs(A, B, &bra);
  build_prim_pairs(Cc, D, &ket);
  std::vector<double> acc(C::kTargets, 0.0);
  g_prims_used = (int)(bra.size() * ket.size());
  if (BOYS == kBoysFarForm) {
    g_min_x = 1e300;
    for (const PrimPair& k : ket)
      for (const PrimPair& b : bra) {
        const double dx = b.Px - k.Px, dy = b.Py - k.Py, dz = b.Pz - k.Pz;
        g_min_x = std::min(g_min_x, b.zeta * k.zeta / (b.zeta + k.zeta) * (dx * dx + dy * dy + dz * dz));
      }
    const PairBound pb = bound_prim_pairs(bra), pk = bound_prim_pairs(ket);
    // the kernels' own test (eri_core.h quartet_regime_f, single precision, conservative rounding)
    const PairBoundF fb = make_pair_bound_f(pb.M[0], pb.M[1], pb.M[2], pb.rad, pb.zmin);
    const PairBoundF fk = make_pair_bound_f(pk.M[0], pk.M[1], pk.M[2], pk.rad, pk.zmin);
    g_proved_far = quartet_regime_f<2>(fb, fk, (float)far_proven_x(C::kL), 0.f, 1) == 0;
  }
  for (const PrimPair& k : ket)
    for (const PrimPair& b : bra)
      if (BOYS == kBoysFarForm)
        primitive_quartet_far<C>(b, k, A.ctr[0], A.ctr[1], A.ctr[2], Cc.ctr[0], Cc.ctr[1],
                                 Cc.ctr[2], acc.data());
      else
        primitive_quartet<C, BOYS == kBoysFarForm ? kBoysExact : BOYS>(
            b, k, A.ctr[0], A.ctr[1], A.ctr[2], Cc.ctr[0], Cc.ctr[1], Cc.ctr[2], tabs, acc.data());
  C::hrr(acc.data(), A.ctr[0] - B.ctr[0], A.ctr[1] - B.ctr[1], A.ctr[2] - B.ctr[2],
         Cc.ctr[0] - D.ctr[0], Cc.ctr[1] - D.ctr[1], Cc.ctr[2] - D.ctr[2], out);
  const int na = ncart(A.l), nb = ncart(B.l), nc = ncart(Cc.l), nd = ncart(D.l);
  for (int a = 0; a < na; ++a)
    for (int b = 0; b < nb; ++b)
      for (int c = 0; c < nc; ++c)
        for (int d = 0; d < nd; ++d)
          out[((a * nb + b) * nc + c) * nd + d] *= ss.compscale[A.l][a] * ss.compscale[B.l][b] *
                                                   ss.compscale[Cc.l][c] * ss.compscale[D.l][d];
}
}  // namespace rchem

using namespace rchem;

struct AllTabs {
  std::vector<double> exact, dthr, ddirect;
  bool delta_ok = false;
  BoysTabs tabs(int L) const {
    BoysTabs t;
    t.exact = exact.data() + (size_t)L * kBoysTableLen;
    t.delta.thr = dthr.data();
    t.delta.direct = ddirect.data();
    return t;
  }
};
static const AllTabs& all_tabs() {
  static AllTabs T;
  if (T.exact.empty()) {
    build_boys_tables(&T.exact);
    T.delta_ok = build_boys_delta_tables(&T.dthr, &T.ddirect);
  }
  return T;
}

static Basis from_flat(int n, const double* origins, const int32_t* powers,
                       const int32_t* prim_offset, const double* exps, const double* coefs,
                       const double* norms) {
  Basis b;
  for (int i = 0; i < n; ++i) {
    CGTO g;
    for (int d = 0; d < 3; ++d) { g.origin[d] = origins[3 * i + d]; g.powers[d] = powers[3 * i + d]; }
    for (int p = prim_offset[i]; p < prim_offset[i + 1]; ++p) {
      PGTO pg;
      for (int d = 0; d < 3; ++d) { pg.origin[d] = g.origin[d]; pg.powers[d] = g.powers[d]; }
      pg.exponent = exps[p];
      pg.norm = norms[p];
      g.primitives.push_back(pg);
      g.coefs.push_back(coefs[p]);
    }
    b.cgtos.push_back(g);
  }
  return b;
}

extern "C" int hostcheck_nshells(int n, const double* origins, const int32_t* powers,
                                 const int32_t* prim_offset, const double* exps,
                                 const double* coefs, const double* norms, int32_t* l_out,
                                 int32_t* bf0_out) {
  Basis b = from_flat(n, origins, powers, prim_offset, exps, coefs, norms);
  ShellSet ss;
  std::string err;
  if (!group_shells(b, &ss, &err, g_fuse_sp != 0)) return -1;
  for (size_t i = 0; i < ss.shells.size(); ++i) {
    if (l_out) l_out[i] = ss.shells[i].l;
    if (bf0_out) bf0_out[i] = ss.shells[i].bf0;
  }
  return (int)ss.shells.size();
}

// Evaluates the Cartesian block of shell quartet (sa sb|sc sd); the shells must already be in
// class order (la>=lb, lc>=ld, (la,lb)>=(lc,ld)).  Returns the number of integrals or <0.
extern "C" int hostcheck_shell_quartet(int n, const double* origins, const int32_t* powers,
                                       const int32_t* prim_offset, const double* exps,
                                       const double* coefs, const double* norms, int sa, int sb,
                                       int sc, int sd, int boys, double* out) {
  Basis b = from_flat(n, origins, powers, prim_offset, exps, coefs, norms);
  ShellSet ss;
  std::string err;
  if (!group_shells(b, &ss, &err, g_fuse_sp != 0)) return -1;
  const AllTabs& T = all_tabs();
  const Shell &A = ss.shells[sa], &B = ss.shells[sb], &C = ss.shells[sc], &D = ss.shells[sd];
#define X(la, lb, lc, ld, tag)                                                              \
  if (A.l == la && B.l == lb && C.l == lc && D.l == ld) {                                   \
    using Cl = EriClass<la, lb, lc, ld>;                                                    \
    if (boys == kBoysReference)                                                             \
      shell_quartet<Cl, kBoysReference>(ss, A, B, C, D, T.tabs(Cl::kL), out);               \
    else if (boys == kBoysFarForm)                                                          \
      shell_quartet<Cl, kBoysFarForm>(ss, A, B, C, D, T.tabs(Cl::kL), out);                 \
    else                                                                                    \
      shell_quartet<Cl, kBoysExact>(ss, A, B, C, D, T.tabs(Cl::kL), out);                   \
    return Cl::kOut;                                                                        \
  }
  RCHEM_ERI_CLASSES(X)
#undef X
  return -2;
}

// what the last far-form call (boys = 4) saw: smallest Boys argument, bounding-sphere verdict
extern "C" int hostcheck_last_prims_used() { return g_prims_used; }
extern "C" void hostcheck_set_fuse(int on) { g_fuse_sp = on; }
extern "C" double hostcheck_last_min_x() { return g_min_x; }
extern "C" int hostcheck_last_proved_far() { return g_proved_far; }
extern "C" double hostcheck_far_threshold(int L) { return far_proven_x(L); }

extern "C" int hostcheck_ref_tables_ok() {
  return all_tabs().delta_ok ? 1 : 0;
}

// boys: 0 (or 3) = reference as exact minus tabulated correction (the path the kernels use),
// 1 = exact, 2 = faithful reference loops
extern "C" void hostcheck_boys(int boys, int L, double x, double* F) {
  const AllTabs& T = all_tabs();
  if (boys == 2) { boys_reference_faithful<8>(x, F); return; }
  if (boys == 3 || boys == kBoysReference) {  // the kernels' path (eri_core.h boys_reference)
    auto exact_x = [&]() { return x; };
    switch (L) {
      case 0: boys_reference<0>(x, T.tabs(0), exact_x, F); break;
      case 2: boys_reference<2>(x, T.tabs(2), exact_x, F); break;
      case 4: boys_reference<4>(x, T.tabs(4), exact_x, F); break;
      default: boys_reference<8>(x, T.tabs(8), exact_x, F); break;
    }
    return;
  }
  {
    switch (L) {
      case 0: boys_exact<0>(x, T.tabs(0).exact, F); break;
      case 2: boys_exact<2>(x, T.tabs(2).exact, F); break;
      case 4: boys_exact<4>(x, T.tabs(4).exact, F); break;
      default: boys_exact<8>(x, T.tabs(8).exact, F); break;
    }
  }
}
