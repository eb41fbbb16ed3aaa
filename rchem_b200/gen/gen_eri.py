#!/usr/bin/env python3
"""Code generator for the per-class ERI recurrences (Head-Gordon--Pople scheme).

For every class (ta tb|tc td) of shell TYPES (s, p, d and the fused sp shell, see SP below)
with ta>=tb, tc>=td, (ta,tb)>=(tc,td) this script emits straight-line C++ (usable from CUDA device code and from a
plain host compiler) for

  * the Obara--Saika vertical recurrence (VRR) that turns the scaled Boys values
    F_0..F_L of ONE primitive quartet into the [e0|f0] targets, la<=|e|<=la+lb,
    lc<=|f|<=lc+ld, and adds them into the contraction accumulators, and
  * the horizontal recurrence (HRR) that turns the contracted [e0|f0] block into the
    Cartesian (ab|cd) block of the shell quartet.

The reference evaluates the same recurrences recursively (chgp.c:44-135 contr_hrr /
contr_vrr, chgp.c:412-586 vrr_recursive); here every intermediate is computed once.
The VRR picks, per target, the Cartesian direction with the fewest terms (2, 3 or 4),
which is the "minimal HGP" operation-count model of SURVEY.md section 8(d).

Component order inside a shell follows shell::get_ijk_list (shell.rs:1-12).

Outputs (into --outdir):
  eri_class_<abcd>.inc   one per class: struct EriClass<la,lb,lc,ld> specialisation
  eri_classes.h          includes + class table
  flops.json             op counts of the emitted code (mul/add = 1, fma = 2)
"""
import argparse
import json
import os
import sys

LMAX = 2
AX = "xyz"

# Shell TYPE codes of a class: 0 = s, 1 = p, 2 = d, 3 = sp -- a FUSED s+p shell (the "L" shells
# of STO-3G / 6-31G: one s and one p contraction on the same exponents, basis.rs:190-201 emits
# them as 4 consecutive functions s, px, py, pz).  A fused shell has two VARIANTS (its s part
# and its p part) with their own contraction coefficients; a quartet of fused shells shares
# every primitive-quartet quantity (Boys values, the whole VRR) between its variants, which
# only differ in the weight a VRR value is accumulated with.
SP = 3
SPLIT_TARGETS = 100  # classes with more contraction accumulators are emitted in parts (gen_class)


def variants(t):
    return [0, 1] if t == SP else [t]


def lmax_of(t):
    return 1 if t == SP else t


def functions(t):
    """functions of a shell of type t in basis order: (variant index, l, Cartesian powers)"""
    return [(vi, l, c) for vi, l in enumerate(variants(t)) for c in cart(l)]


def cart(l):
    """shell::get_ijk_list (shell.rs:1-12)"""
    return [(l + 1 - a, a - b, b - 1) for a in range(1, l + 2) for b in range(1, a + 1)]


def dec(t, i):
    return tuple(v - (1 if k == i else 0) for k, v in enumerate(t))


def inc(t, i):
    return tuple(v + (1 if k == i else 0) for k, v in enumerate(t))


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.flops = 0

    def tmp(self, expr, flops):
        name = f"t{self.n}"
        self.n += 1
        self.lines.append(f"const double {name} = {expr};")
        self.flops += flops
        return name


class VRR:
    """Memoised OS vertical recurrence for [e0|f0]^(m)."""

    def __init__(self, em):
        self.em = em
        self.memo = {}
        self.gmemo = {}

    def val(self, e, f, m):
        key = (e, f, m)
        if key in self.memo:
            return self.memo[key]
        if sum(e) == 0 and sum(f) == 0:
            r = f"F[{m}]"
        elif sum(f) > 0:
            r = self._reduce(e, f, m, ket=True)
        else:
            r = self._reduce(e, f, m, ket=False)
        self.memo[key] = r
        return r

    def gdiff(self, e, f, m, ket):
        """[e|f]^(m) - (rho/eta or rho/zeta) [e|f]^(m+1), shared between targets"""
        key = (e, f, m, ket)
        if key in self.gmemo:
            return self.gmemo[key]
        a = self.val(e, f, m)
        b = self.val(e, f, m + 1)
        fac = "roe" if ket else "roz"
        r = self.em.tmp(f"fma(-{fac}, {b}, {a})", 2)
        self.gmemo[key] = r
        return r

    def _reduce(self, e, f, m, ket):
        tgt = f if ket else e
        # choose the direction with the fewest terms
        best = None
        for i in range(3):
            if tgt[i] == 0:
                continue
            low = dec(tgt, i)
            cost = 2 + (1 if low[i] > 0 else 0)
            if ket and e[i] > 0:
                cost += 1
            # tie-break: prefer directions whose lower terms are likely shared (larger i last)
            if best is None or cost < best[0]:
                best = (cost, i)
        i = best[1]
        low = dec(tgt, i)
        if ket:
            e0, f0 = e, low
            c1, c2 = f"QC{AX[i]}", f"WQ{AX[i]}"
        else:
            e0, f0 = low, f
            c1, c2 = f"PA{AX[i]}", f"WP{AX[i]}"
        v0 = self.val(e0, f0, m)
        v1 = self.val(e0, f0, m + 1)
        expr = f"fma({c1}, {v0}, {c2} * {v1})"
        flops = 3
        if low[i] > 0:
            if ket:
                g = self.gdiff(e0, dec(f0, i), m, True)
                coef = f"oo2e" if low[i] == 1 else f"({low[i]}.0 * oo2e)"
            else:
                g = self.gdiff(dec(e0, i), f0, m, False)
                coef = f"oo2z" if low[i] == 1 else f"({low[i]}.0 * oo2z)"
            expr = f"fma({coef}, {g}, {expr})"
            flops += 2
        if ket and e[i] > 0:
            v = self.val(dec(e, i), f0, m + 1)
            coef = f"oo2ze" if e[i] == 1 else f"({e[i]}.0 * oo2ze)"
            expr = f"fma({coef}, {v}, {expr})"
            flops += 2
        return self.em.tmp(expr, flops)


def gen_class(ta, tb, tc, td, force_mono=False):
    L = lmax_of(ta) + lmax_of(tb) + lmax_of(tc) + lmax_of(td)
    # variants of the bra / ket shell pair: index = ia * nvar(B) + ib
    bra_vars = [(la, lb) for la in variants(ta) for lb in variants(tb)]
    ket_vars = [(lc, ld) for lc in variants(tc) for ld in variants(td)]
    nvb, nvk = len(bra_vars), len(ket_vars)
    fused = nvb * nvk > 1
    tindex = {}
    for vb, (la, lb) in enumerate(bra_vars):
        for vk, (lc, ld) in enumerate(ket_vars):
            for l in range(la, la + lb + 1):
                for e in cart(l):
                    for lf in range(lc, lc + ld + 1):
                        for f in cart(lf):
                            tindex[(vb, vk, e, f)] = len(tindex)
    nt = len(tindex)

    # ---- VRR ---------------------------------------------------------------
    # Classes past SPLIT_TARGETS accumulators cannot live in registers: one straight-line function
    # of several thousand values sends ptxas into its spill-everything fallback (32 registers,
    # 30-70 kB of stack, FP64 pipe ~1 % busy on (dd|dd)).  Those classes are emitted in PARTS --
    # one small __noinline__ function per bra component e (all ket targets of that e), each
    # with its own memo table -- so every part is register-allocated on its own; the lower
    # [e'0|f'0] values the parts share are recomputed (1.6-1.9x the flops of the monolithic form).
    split = nt > SPLIT_TARGETS and not force_mono
    groups = {}
    for key in tindex:
        groups.setdefault(key[2] if split else None, []).append(key)
    vrr_parts = []
    vrr_flops = 0
    for keys in groups.values():
        em = Emitter()
        v = VRR(em)
        acc_lines = []
        for (vb, vk, e, f) in keys:
            idx = tindex[(vb, vk, e, f)]
            name = v.val(e, f, 0)
            if fused:
                acc_lines.append(f"acc[{idx}] = fma(W[{vb * nvk + vk}], {name}, acc[{idx}]);")
            else:
                acc_lines.append(f"acc[{idx}] += {name};")
        vrr_flops += em.flops + len(keys) * (2 if fused else 1)  # one add (fma) per target for the contraction
        vrr_parts.append(em.lines + acc_lines)

    # ---- HRR ---------------------------------------------------------------
    # (split classes: one part per bra function a, again with fresh memo tables)
    st = {"hm": Emitter(), "bmemo": {}, "kmemo": {}}

    def hb(vb, vk, a, b, f):
        key = (vb, vk, a, b, f)
        if key in st["bmemo"]:
            return st["bmemo"][key]
        if sum(b) == 0:
            r = f"acc[{tindex[(vb, vk, a, f)]}]"
        else:
            i = next(k for k in range(3) if b[k] > 0)
            hi = hb(vb, vk, inc(a, i), dec(b, i), f)
            lo = hb(vb, vk, a, dec(b, i), f)
            r = st["hm"].tmp(f"fma(AB{AX[i]}, {lo}, {hi})", 2)
        st["bmemo"][key] = r
        return r

    def hk(vb, vk, a, b, c, d):
        key = (vb, vk, a, b, c, d)
        if key in st["kmemo"]:
            return st["kmemo"][key]
        if sum(d) == 0:
            r = hb(vb, vk, a, b, c)
        else:
            i = next(k for k in range(3) if d[k] > 0)
            hi = hk(vb, vk, a, b, inc(c, i), dec(d, i))
            lo = hk(vb, vk, a, b, c, dec(d, i))
            r = st["hm"].tmp(f"fma(CD{AX[i]}, {lo}, {hi})", 2)
        st["kmemo"][key] = r
        return r

    fa, fb, fc, fd = functions(ta), functions(tb), functions(tc), functions(td)
    nout = len(fa) * len(fb) * len(fc) * len(fd)
    hrr_parts = []
    hrr_flops = 0
    o = 0
    all_out = []
    for (ia, _, a) in fa:
        out_lines = []
        if split:
            st = {"hm": Emitter(), "bmemo": {}, "kmemo": {}}
        for (ib, _, b) in fb:
            for (ic, _, c) in fc:
                for (id_, _, d) in fd:
                    vb = ia * len(variants(tb)) + ib
                    vk = ic * len(variants(td)) + id_
                    out_lines.append(f"out[{o}] = {hk(vb, vk, a, b, c, d)};")
                    o += 1
        if split:
            hrr_flops += st["hm"].flops
            hrr_parts.append(st["hm"].lines + out_lines)
        else:
            all_out += out_lines
    if not split:
        hrr_flops = st["hm"].flops
        hrr_parts.append(st["hm"].lines + all_out)

    tag = f"{ta}{tb}{tc}{td}"
    names = ["s", "p", "d", "sp"]
    src = []
    src.append(f"// generated by rchem_b200/gen/gen_eri.py -- do not edit")
    src.append(f"// class ({names[ta]} {names[tb]}|{names[tc]} {names[td]}): {nt} contraction accumulators, "
               f"{nout} integrals, emitted flops/primitive {vrr_flops}, HRR flops {hrr_flops}")
    src.append(f"template <> struct EriClass<{ta}, {tb}, {tc}, {td}> {{")
    src.append(f"  static constexpr int kL = {L};")
    src.append(f"  static constexpr int kTargets = {nt};")
    src.append(f"  static constexpr int kOut = {nout};")
    src.append(f"  static constexpr int kNVb = {nvb};  // weight variants of the bra / ket shell pair")
    src.append(f"  static constexpr int kNVk = {nvk};")
    src.append(f"  static constexpr int kVrrFlops = {vrr_flops};")
    src.append(f"  static constexpr int kHrrFlops = {hrr_flops};")
    geom_names = ["PAx", "PAy", "PAz", "WPx", "WPy", "WPz", "QCx", "QCy", "QCz", "WQx", "WQy", "WQz",
                  "oo2z", "oo2e", "oo2ze", "roz", "roe"]
    wdecl = "const double* __restrict__ W, " if fused else ""
    warg = "W, " if fused else ""

    def emit_vrr(name, body, qual):
        src.append(f"  {qual} static void {name}(const double* __restrict__ F, const VrrGeom& g, "
                   f"{wdecl}double* __restrict__ acc) {{")
        used = "\n".join(body)
        for nm in geom_names:
            if nm in used:
                src.append(f"    const double {nm} = g.{nm};")
        src.extend("    " + l for l in body)
        src.append("  }")

    def emit_hrr(name, body, qual):
        src.append(f"  {qual} static void {name}(const double* __restrict__ acc, double ABx, double ABy, "
                   "double ABz, double CDx, double CDy, double CDz, double* __restrict__ out) {")
        src.append("    (void)ABx; (void)ABy; (void)ABz; (void)CDx; (void)CDy; (void)CDz;")
        src.extend("    " + l for l in body)
        src.append("  }")

    if fused:
        src.append("  // W[vb * kNVk + vk]: contraction weight of (bra variant vb, ket variant vk)")
    if not split:
        emit_vrr("vrr", vrr_parts[0], "RCHEM_HD")
        emit_hrr("hrr", hrr_parts[0], "RCHEM_HD")
    else:
        src.append(f"  // emitted in {len(vrr_parts)} VRR parts (one per bra component) and {len(hrr_parts)} HRR parts "
                   "(one per bra function): see gen_eri.py")
        for k, body in enumerate(vrr_parts):
            emit_vrr(f"vrr_part{k}", body, "RCHEM_PART")
        emit_vrr("vrr", [f"vrr_part{k}(F, g, {warg}acc);" for k in range(len(vrr_parts))], "RCHEM_HD")
        for k, body in enumerate(hrr_parts):
            emit_hrr(f"hrr_part{k}", body, "RCHEM_PART")
        emit_hrr("hrr", [f"hrr_part{k}(acc, ABx, ABy, ABz, CDx, CDy, CDz, out);" for k in range(len(hrr_parts))],
                 "RCHEM_HD")
    src.append("};")
    info = dict(targets=nt, out=nout, vrr_flops=vrr_flops, hrr_flops=hrr_flops)
    if split:
        # the op-count MODEL is the monolithic form's; the parts recompute shared values
        mono = gen_class(ta, tb, tc, td, force_mono=True)[2]
        info = dict(targets=nt, out=nout, vrr_flops=mono["vrr_flops"], hrr_flops=mono["hrr_flops"],
                    emitted_vrr_flops=vrr_flops, emitted_hrr_flops=hrr_flops,
                    vrr_parts=len(vrr_parts), hrr_parts=len(hrr_parts))
    return tag, "\n".join(src) + "\n", info


def classes(lmax=LMAX):
    """Classes over the type sets {s,p,d} (segmented shells) and {s,sp} (bases made of s and
    fused sp shells only: STO-3G, 6-31G): (ta>=tb | tc>=td), bra pair >= ket pair in the order
    pair(ta,tb) = ta(ta+1)/2 + tb.  (Classes mixing sp and d shells would need up to 899
    contraction accumulators per thread; a basis with d shells keeps its s and p shells
    segmented.)"""
    out = []
    for types in ([t for t in (0, 1, 2) if t <= lmax], [0, SP] if lmax >= 1 else []):
        pairs = sorted(((a, b) for a in types for b in types if a >= b), key=lambda p: p[0] * (p[0] + 1) // 2 + p[1])
        for i, (ta, tb) in enumerate(pairs):
            for (tc, td) in pairs[: i + 1]:
                if (ta, tb, tc, td) not in out:
                    out.append((ta, tb, tc, td))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--outdir", default=os.path.join(os.path.dirname(__file__), "..", "csrc", "gen"))
    ap.add_argument("--lmax", type=int, default=LMAX,
                    help="highest shell angular momentum (2 = s,p,d: 21 classes; 1 is a fast dev build)")
    args = ap.parse_args()
    os.makedirs(args.outdir, exist_ok=True)
    table = {}
    tags = []
    for (la, lb, lc, ld) in classes(args.lmax):
        tag, src, info = gen_class(la, lb, lc, ld)
        with open(os.path.join(args.outdir, f"eri_class_{tag}.inc"), "w") as fh:
            fh.write(src)
        table[tag] = info
        tags.append((la, lb, lc, ld, tag))
    with open(os.path.join(args.outdir, "flops.json"), "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    with open(os.path.join(args.outdir, "eri_class_list.h"), "w") as fh:
        fh.write("// generated by rchem_b200/gen/gen_eri.py -- do not edit\n")
        fh.write("// X(la, lb, lc, ld, tag)\n#define RCHEM_ERI_CLASSES(X) \\\n")
        fh.write(" \\\n".join(f"  X({la}, {lb}, {lc}, {ld}, {tag})" for la, lb, lc, ld, tag in tags))
        fh.write("\n")
    print(f"generated {len(tags)} classes into {os.path.abspath(args.outdir)}", file=sys.stderr)


if __name__ == "__main__":
    main()
