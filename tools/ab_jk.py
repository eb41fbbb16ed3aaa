"""A/B timing of direct J/K builds across library builds and tuning knobs (development tool).

    python tools/ab_jk.py [nw basis tau]        default: 96 6-31G 1e-10 (the headline workload)

Every configuration runs in its own process (the knobs are read once per process):
  base   rchem_b200/librchem_b200_base.so if present (an older build, RCHEM_B200_LIB)
  new    the in-tree library with RCHEM_FAR / RCHEM_LIGHT = 0/1
and reports the best of 4 kernel times for both Boys flavours plus max|J - J_first|, max|K - K_first|
against the first configuration that ran (reference flavour)."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os, numpy as np
sys.path.insert(0, %r)
import rchem_b200 as rc
from rchem_b200 import geometry as geo
nw, bas, tau, tag = int(sys.argv[1]), sys.argv[2], float(sys.argv[3]), sys.argv[4]
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas); b.set_schwarz_tau(tau)
n = b.nbf; D = geo.synthetic_density(n)
res = {"tag": tag, "n": n}
for boys in (0, 1):
    b.set_boys(boys)
    J, K = np.zeros((n, n)), np.zeros((n, n))
    best = 1e9
    for _ in range(5):
        rc.JK_direct(J, K, b, D); best = min(best, b.stats()["kernel_ms"])
    res["ms_boys%%d" %% boys] = best
    ref = "/tmp/ab_jk_ref_%%d_%%s_%%d.npz" %% (nw, bas.replace("*", "s"), boys)
    if not os.path.exists(ref):
        np.savez(ref, J=J, K=K)
    else:
        g = np.load(ref)
        res["dJ_boys%%d" %% boys] = float(np.abs(J - g["J"]).max()); res["dK_boys%%d" %% boys] = float(np.abs(K - g["K"]).max())
s = b.stats(); res["quartets"] = s["shell_quartets"]; res["launches"] = s["launches"]
print("RESULT " + json.dumps(res))
''' % ROOT

def main():
    nw, bas, tau = (sys.argv[1:4] + ["96", "6-31G", "1e-10"][len(sys.argv) - 1:])[:3]
    base = os.path.join(ROOT, "rchem_b200", "librchem_b200_base.so")
    cfgs = []
    if os.path.exists(base) and "*" not in bas and not os.environ.get("AB_NOBASE"):
        cfgs.append(("base", {"RCHEM_B200_LIB": base}))
    combos = os.environ.get("AB_COMBOS", "00,01,10,11").split(",")  # far,light digits
    for far, light in [c for c in combos if c]:
        cfgs.append((f"new far={far} light={light}", {"RCHEM_FAR": far, "RCHEM_LIGHT": light}))
    # AB_VARIANTS="tag|library file under rchem_b200/ (or empty)|ENV=VAL,ENV=VAL;..."
    for spec in [v for v in os.environ.get("AB_VARIANTS", "").split(";") if v]:
        tag, lib, envs = (spec.split("|") + ["", ""])[:3]
        env = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
        if lib:
            env["RCHEM_B200_LIB"] = os.path.join(ROOT, "rchem_b200", lib)
        cfgs.append((tag, env))
    for tag, env in cfgs:
        e = dict(os.environ); e.update(env)
        p = subprocess.run([sys.executable, "-c", CHILD, nw, bas, tau, tag], env=e, capture_output=True, text=True, timeout=900)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(tag, "FAILED", p.stdout[-500:], p.stderr[-1500:]); continue
        r = json.loads(line[0][7:])
        print(f"(H2O){nw} {bas} tau={tau} N={r['n']} {tag:24s} ref {r['ms_boys0']:8.2f} ms  exact {r['ms_boys1']:8.2f} ms  "
              f"dJ/dK ref {r.get('dJ_boys0', 0):.1e}/{r.get('dK_boys0', 0):.1e} exact {r.get('dJ_boys1', 0):.1e}/{r.get('dK_boys1', 0):.1e} "
              f"quartets {r['quartets']:.3e} launches {r['launches']}", flush=True)

if __name__ == "__main__":
    main()
