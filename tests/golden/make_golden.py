#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference C sources.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
It compiles oracle/_ref/libpyquante2_ref.so (oracle/Makefile) and calls the reference's own
coulomb_repulsion / hrr / Fgamma / ijkl2intindex; the loop nests of basis.rs (which cannot be
run here: no Rust toolchain) are the oracle's restatement driven by that reference kernel.
Outputs: tests/golden/*.npz (small; committed)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from rchem_b200 import geometry as geo  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    orc.build(force=True)
    ref = orc.ref_lib()
    assert ref is not None, "reference checkout not available"
    orc.use_reference_kernel(True)

    # 1. Fgamma grid (cints.c:302-308)
    ms = np.arange(0, 13, dtype=np.float64)
    xs = np.concatenate([[0.0, 1e-12, 1e-9, 1e-8, 2e-8], np.linspace(1e-4, 60.0, 400),
                         ms + 1.5 - 1e-9, ms + 1.5 + 1e-9])
    fg = np.array([[ref.Fgamma(m, x) for x in xs] for m in ms])
    np.savez(os.path.join(OUT, "fgamma_ref.npz"), m=ms, x=xs, F=fg)

    # 2. random primitive integrals up to f functions (coulomb_repulsion, hrr)
    rng = np.random.default_rng(7)
    n = 400
    centres = rng.uniform(-1.5, 1.5, size=(n, 12))
    centres[:40, 3:6] = centres[:40, 0:3]      # some two-centre / one-centre cases
    centres[:10, 6:] = np.tile(centres[:10, 0:3], 2)
    alphas = np.exp(rng.uniform(np.log(0.15), np.log(40.0), size=(n, 4)))
    norms = rng.uniform(0.5, 1.5, size=(n, 4))
    powers = np.zeros((n, 12), dtype=np.int32)
    for t in range(n):
        for c in range(4):
            L = rng.integers(0, 4 if t % 4 == 0 else 3)
            for _ in range(L):
                powers[t, 3 * c + rng.integers(0, 3)] += 1
    tho = np.array([orc.ref_coulomb_repulsion(centres[t, 0:3], centres[t, 3:6], centres[t, 6:9],
                                              centres[t, 9:12], powers[t], alphas[t], norms[t])
                    for t in range(n)])
    hgp = np.array([orc.ref_coulomb_repulsion(centres[t, 0:3], centres[t, 3:6], centres[t, 6:9],
                                              centres[t, 9:12], powers[t], alphas[t], norms[t],
                                              fn="hrr") for t in range(n)])
    np.savez(os.path.join(OUT, "primitives_ref.npz"), centres=centres, alphas=alphas, norms=norms,
             powers=powers, tho=tho, hgp=hgp)

    # 3. water_crawford and water.xyz, STO-3G: full tensor, J/K (configs 1 and 2 of BASELINE.json)
    for tag, spec in (("water_crawford", geo.WATER_CRAWFORD), ("water", geo.WATER)):
        z, x = geo.molecule(spec)
        b = orc.make_basis(z, x, "STO-3G")
        I = orc.build_I(b)
        D = geo.synthetic_density(b.n)
        J, K = orc.jk_direct(b, D)
        np.savez(os.path.join(OUT, f"{tag}_sto3g.npz"), atomnos=z, coords=x, I=I, D=D, J=J, K=K)

    # 4. water_crawford 6-31G* (s, p, d): full tensor is 19^4 = 130k doubles -> 1 MB; keep a
    #    seeded sample of 6000 elements plus J/K
    z, x = geo.molecule(geo.WATER_CRAWFORD)
    b = orc.make_basis(z, x, "6-31G*")
    I = orc.build_I(b)
    D = geo.synthetic_density(b.n)
    J, K = orc.jk_inmem(I, D)
    idx = np.random.default_rng(11).integers(0, b.n, size=(6000, 4)).astype(np.int32)
    np.savez(os.path.join(OUT, "water_crawford_631gs.npz"), atomnos=z, coords=x, idx=idx,
             vals=I[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]], D=D, J=J, K=K)

    # 5. (H2O)_2 6-31G*: sampled contracted integrals (multi-centre d functions)
    z, x = geo.water_cluster(2)
    b = orc.make_basis(z, x, "6-31G*")
    idx = np.random.default_rng(13).integers(0, b.n, size=(4000, 4)).astype(np.int32)
    vals = orc.eval_quartets(b, idx)
    np.savez(os.path.join(OUT, "water2_631gs_sample.npz"), atomnos=z, coords=x, idx=idx, vals=vals)

    # 6. canonical index (cints.c:375-387), within the reference's 32-bit range
    ijkl = np.random.default_rng(17).integers(0, 200, size=(500, 4)).astype(np.int64)
    can = np.array([ref.ijkl2intindex(*map(int, q)) for q in ijkl], dtype=np.int64)
    np.savez(os.path.join(OUT, "ijkl_ref.npz"), ijkl=ijkl, index=can)
    orc.use_reference_kernel(False)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
