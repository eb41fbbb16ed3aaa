"""Distribution of surviving-ket counts nq over (task, bra pair): where do the quartets sit?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo
nw, bas, tau = int(sys.argv[1]), sys.argv[2], float(sys.argv[3])
z, x = geo.water_cluster(nw)
b = rc.Basis.new(z, x, bas)
sa, sb, batch, Q = b.schwarz()
bins = [0, 1, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 1 << 30]
hist_pairs = np.zeros(len(bins) - 1); hist_q = np.zeros(len(bins) - 1); waste = 0.0; tot = 0.0
for bi in np.unique(batch):
    Qb = Q[batch == bi]
    for ki in np.unique(batch):
        if ki > bi: continue
        Qk = Q[batch == ki]
        # nq[p] = #{q: Qb[p]*Qk[q] >= tau}, Qk descending
        nq = np.searchsorted(-Qk, -(tau / np.maximum(Qb, 1e-300)), side="right")
        if bi == ki: nq = np.minimum(nq, np.arange(len(Qb)) + 1)
        h, _ = np.histogram(nq, bins=bins); hist_pairs += h
        hq, _ = np.histogram(nq, bins=bins, weights=nq); hist_q += hq
        heavy = nq >= 512
        nb = np.ceil(nq[heavy] / 4096)
        per = np.ceil(np.ceil(nq[heavy] / nb) / 32) * 32
        passes = np.ceil(per / 512)
        waste += (nb * passes * 512 - nq[heavy]).sum(); tot += nq[heavy].sum()
print("nq bin        pairs        quartets   share")
for i in range(len(bins) - 1):
    print(f"[{bins[i]:5d},{bins[i+1]:10d}) {hist_pairs[i]:10.0f} {hist_q[i]:14.4g} {100*hist_q[i]/hist_q.sum():6.1f}%")
print(f"heavy quartets {tot:.4g}, idle thread-slots from partial passes {waste:.4g} ({100*waste/(tot+waste):.1f}% of slots)")
