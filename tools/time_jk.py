"""Quick timing of direct J/K builds: python tools/time_jk.py nw basis tau boys reps"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo
nw, bas, tau, boys, reps = int(sys.argv[1]), sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas)
b.set_schwarz_tau(tau); b.set_boys(boys)
n = b.nbf
D = geo.synthetic_density(n)
J, K = np.zeros((n, n)), np.zeros((n, n))
t = time.time(); rc.JK_direct(J, K, b, D); t0 = time.time() - t
best = 1e9
for _ in range(reps):
    rc.JK_direct(J, K, b, D); best = min(best, b.stats()["kernel_ms"])
s = b.stats()
print(f"(H2O){nw} {bas} N={n} tau={tau} boys={boys}: first {t0:.2f}s kernel {best:.2f} ms  quartets {s['shell_quartets']:.3e}/{s['shell_quartets_all']:.3e} prim {s['prim_quartets']:.3e} model {s['model_flops']/best/1e9:.2f} TFLOP/s  {s['shell_quartets']/best*1e3:.3e} q/s")
