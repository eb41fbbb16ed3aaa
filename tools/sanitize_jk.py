"""compute-sanitizer driver: one small J/K build through each kernel family (block / light /
chunk forced with RCHEM_OPT_HEAVY_PASSES / RCHEM_OPT_LIGHT_KERNEL), both Boys flavours, far-field
routing on, plus a dense tensor build; results checked against the first configuration.

    compute-sanitizer --tool memcheck  python tools/sanitize_jk.py
    compute-sanitizer --tool racecheck python tools/sanitize_jk.py
"""
import sys
import numpy as np
sys.path.insert(0, ".")
import rchem_b200 as rc
from rchem_b200 import geometry as geo

nw = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bas = sys.argv[2] if len(sys.argv) > 2 else "6-31G"
z, x = geo.water_cluster(nw)
ref = None
for kernel, (passes, light) in {"block": (0.0, True), "light": (1e9, True), "chunk": (1e9, False)}.items():
    for boys in (0, 1):
        b = rc.Basis.new(z, x, bas)
        b.set_boys(boys)
        b.set_heavy_passes(passes)
        b.set_light_kernel(light)
        n = b.nbf
        D = geo.synthetic_density(n)
        J, K = np.zeros((n, n)), np.zeros((n, n))
        rc.JK_direct(J, K, b, D)
        if boys == 0:
            if ref is None:
                ref = (J.copy(), K.copy())
            err = max(np.abs(J - ref[0]).max(), np.abs(K - ref[1]).max())
            assert err < 1e-13, (kernel, err)
        print(kernel, "boys", boys, "ok", b.stats()["launches"], "launches", flush=True)
b = rc.Basis.new(*geo.water_cluster(2), "STO-3G")
I = rc.build_I(b)
print("tensor ok", I.shape, flush=True)
