"""Small driver for ncu: a few J/K builds of one workload (no timing claims)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo
nw, bas, tau, boys, reps = int(sys.argv[1]), sys.argv[2], float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas)
b.set_schwarz_tau(tau)
b.set_boys(boys)
n = b.nbf
D = geo.synthetic_density(n)
J, K = np.zeros((n, n)), np.zeros((n, n))
for _ in range(reps):
    rc.JK_direct(J, K, b, D)
print(b.stats())
