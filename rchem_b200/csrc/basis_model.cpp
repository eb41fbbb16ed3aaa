// basis_model.cpp -- Basis::new equivalent, embedded basis tables, shell grouping.
// See basis_model.h for the reference correspondence.
#include "basis_model.h"

#include <algorithm>
#include <cctype>
#include <map>

namespace rchem {
namespace {

std::string canon(const std::string& s) {
  std::string r;
  for (char c : s)
    if (!std::isspace((unsigned char)c)) r.push_back((char)std::toupper((unsigned char)c));
  if (r == "6-31G(D)") r = "6-31G*";
  return r;
}

using Table = std::map<unsigned, std::vector<BseShell>>;

// STO-3G (Hehre, Stewart, Pople 1969), values as distributed by the Basis Set Exchange;
// these are the numbers SURVEY section 8(d) lists and that reproduce Crawford's water energies.
const Table& sto3g() {
  static const Table t = {
      {1, {{{0}, {3.42525091, 0.62391373, 0.16885540}, {{0.15432897, 0.53532814, 0.44463454}}}}},
      {8,
       {{{0}, {130.70932, 23.808861, 6.4436083}, {{0.15432897, 0.53532814, 0.44463454}}},
        {{0, 1},
         {5.0331513, 1.1695961, 0.3803890},
         {{-0.09996723, 0.39951283, 0.70011547}, {0.15591627, 0.60768372, 0.39195739}}}}},
  };
  return t;
}

// 6-31G (Hehre, Ditchfield, Pople 1972; H from Ditchfield, Hehre, Pople 1971)
const Table& b631g() {
  static const Table t = {
      {1,
       {{{0}, {18.7311370, 2.8253937, 0.6401217}, {{0.03349460, 0.23472695, 0.81375733}}},
        {{0}, {0.1612778}, {{1.0}}}}},
      {8,
       {{{0},
         {5484.6717000, 825.2349500, 188.0469600, 52.9645000, 16.8975700, 5.7996353},
         {{0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.4701930, 0.3585209}}},
        {{0, 1},
         {15.5396160, 3.5999336, 1.0137618},
         {{-0.1107775, -0.1480263, 1.1307670}, {0.0708743, 0.3397528, 0.7271586}}},
        {{0, 1}, {0.2700058}, {{1.0}, {1.0}}}}},
  };
  return t;
}

// 6-31G* = 6-31G + one Cartesian d polarisation function on heavy atoms (Hariharan, Pople 1973)
const Table& b631gs() {
  static const Table t = [] {
    Table u = b631g();
    u[8].push_back({{2}, {0.8}, {{1.0}}});
    return u;
  }();
  return t;
}

}  // namespace

const std::vector<BseShell>* embedded_basis(const std::string& name, unsigned atomno) {
  const std::string c = canon(name);
  const Table* t = nullptr;
  if (c == "STO-3G") t = &sto3g();
  else if (c == "6-31G") t = &b631g();
  else if (c == "6-31G*") t = &b631gs();
  if (!t) return nullptr;
  auto it = t->find(atomno);
  return it == t->end() ? nullptr : &it->second;
}

bool basis_new(const std::vector<uint64_t>& atomnos, const double* coords,
               const std::string& basis_set_name, Basis* out, std::string* err) {
  out->name = basis_set_name;
  out->cgtos.clear();
  for (size_t i = 0; i < atomnos.size(); ++i) {
    const std::vector<BseShell>* element = embedded_basis(basis_set_name, (unsigned)atomnos[i]);
    if (!element) {
      *err = "basis '" + basis_set_name + "' is not embedded for Z=" + std::to_string(atomnos[i]);
      return false;
    }
    const double* xyz = coords + 3 * i;
    for (const BseShell& shell : *element) {
      for (size_t k = 0; k < shell.angular_momentum.size(); ++k) {
        const int am = shell.angular_momentum[k];
        // The reference indexes coefficients by the VALUE of am (basis.rs:193,201), which
        // only works for s / sp shells.  Indexing by position is identical for
        // those and also correct for lone p/d shells.
        const std::vector<double>& coefs = shell.coefficients[k];
        for (const auto& powers : get_ijk_list(am)) {
          CGTO cg;
          for (int d = 0; d < 3; ++d) { cg.origin[d] = xyz[d]; cg.powers[d] = powers[d]; }
          cg.coefs = coefs;
          for (double e : shell.exponents) {
            PGTO pg;
            for (int d = 0; d < 3; ++d) { pg.origin[d] = xyz[d]; pg.powers[d] = powers[d]; }
            pg.exponent = e;
            pg.norm = pgto_normalization(pg.powers, e);
            cg.primitives.push_back(pg);
          }
          out->cgtos.push_back(std::move(cg));
        }
      }
    }
  }
  return true;
}

bool group_shells(const Basis& basis, ShellSet* out, std::string* err, bool fuse_sp) {
  out->shells.clear();
  out->lmax = 0;
  out->fused = false;
  bool have_scale[5] = {false, false, false, false, false};
  const int n = (int)basis.cgtos.size();
  int i = 0;
  while (i < n) {
    const CGTO& first = basis.cgtos[i];
    const int l = first.powers[0] + first.powers[1] + first.powers[2];
    if (l > 4) { *err = "angular momentum > 4 is not supported"; return false; }
    const auto comps = get_ijk_list(l);
    const int nc = (int)comps.size();
    const size_t np = first.primitives.size();
    if (np == 0 || first.coefs.size() != np) { *err = "CGTO without primitives"; return false; }
    if (i + nc > n) { *err = "incomplete Cartesian shell at function " + std::to_string(i); return false; }
    double scale[15];
    for (int k = 0; k < nc; ++k) {
      const CGTO& g = basis.cgtos[i + k];
      bool ok = g.primitives.size() == np && g.coefs.size() == np;
      for (int d = 0; d < 3 && ok; ++d)
        ok = g.powers[d] == comps[k][d] && g.origin[d] == first.origin[d];
      for (size_t p = 0; p < np && ok; ++p)
        ok = g.primitives[p].exponent == first.primitives[p].exponent && g.coefs[p] == first.coefs[p];
      if (!ok) {
        *err = "functions " + std::to_string(i) + ".." + std::to_string(i + nc - 1) +
               " do not form a Cartesian shell in get_ijk_list order";
        return false;
      }
      scale[k] = g.primitives[0].norm / first.primitives[0].norm;
      for (size_t p = 1; p < np; ++p) {
        const double s = g.primitives[p].norm / first.primitives[p].norm;
        if (std::fabs(s - scale[k]) > 1e-13 * std::fabs(scale[k])) {
          *err = "norms of shell at function " + std::to_string(i) + " do not factorise";
          return false;
        }
      }
    }
    if (!have_scale[l]) {
      for (int k = 0; k < nc; ++k) out->compscale[l][k] = scale[k];
      have_scale[l] = true;
    } else {
      for (int k = 0; k < nc; ++k)
        if (std::fabs(out->compscale[l][k] - scale[k]) > 1e-13 * std::fabs(scale[k])) {
          *err = "component norm ratios differ between shells of l=" + std::to_string(l);
          return false;
        }
    }
    Shell sh;
    sh.l = l;
    sh.bf0 = i;
    for (int d = 0; d < 3; ++d) sh.ctr[d] = first.origin[d];
    for (size_t p = 0; p < np; ++p) {
      sh.exps.push_back(first.primitives[p].exponent);
      sh.cn.push_back(first.coefs[p] * first.primitives[p].norm);
    }
    out->shells.push_back(std::move(sh));
    out->lmax = std::max(out->lmax, l);
    i += nc;
  }
  if (fuse_sp && out->lmax == 1) {
    // every p shell must pair with the s shell right before it (same centre, same exponents,
    // consecutive functions): s, px, py, pz of one Basis Set Exchange sp shell
    std::vector<Shell> fusedv;
    bool ok = true;
    const std::vector<Shell>& sh = out->shells;
    for (size_t k = 0; k < sh.size() && ok; ++k) {
      if (sh[k].l == 1) { ok = false; break; }  // a p shell that did not follow its s partner
      if (k + 1 < sh.size() && sh[k + 1].l == 1) {
        const Shell &S = sh[k], &P = sh[k + 1];
        bool same = P.bf0 == S.bf0 + 1 && S.exps.size() == P.exps.size();
        for (int d = 0; d < 3 && same; ++d) same = S.ctr[d] == P.ctr[d];
        for (size_t q = 0; q < S.exps.size() && same; ++q) same = S.exps[q] == P.exps[q];
        if (!same) { ok = false; break; }
        Shell f = S;
        f.l = 3;  // kTypeSP
        f.cn2 = P.cn;
        fusedv.push_back(std::move(f));
        ++k;
      } else {
        fusedv.push_back(sh[k]);
      }
    }
    if (ok) {
      out->shells.swap(fusedv);
      out->fused = true;
      for (int k = 0; k < 15; ++k) out->compscale[3][k] = 1.0;
    }
  }
  return true;
}

}  // namespace rchem
