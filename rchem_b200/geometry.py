"""Geometry helpers for the harness: xyz reader and the synthetic water-cluster generator of
SURVEY.md section 8(d).  Coordinates are used RAW (no Angstrom->bohr conversion), like the
reference driver (rchem.rs:29-35)."""
import math

import numpy as np

ATOMIC_NUMBERS = {"H": 1, "He": 2, "C": 6, "N": 7, "O": 8}

# water_crawford.xyz of the reference (bohr)
WATER_CRAWFORD = (
    (8, (0.000000000000, -0.143225816552, 0.000000000000)),
    (1, (1.638036840407, 1.136548822547, -0.000000000000)),
    (1, (-1.638036840407, 1.136548822547, -0.000000000000)),
)
# water.xyz of the reference
WATER = (
    (8, (0.000000000000, 0.000000000000, -0.075791843589)),
    (1, (0.000000000000, -0.866811828967, 0.601435779270)),
    (1, (0.000000000000, 0.866811828967, 0.601435779270)),
)


def read_xyz(path):
    with open(path) as fh:
        lines = fh.read().splitlines()
    n = int(lines[0].split()[0])
    atomnos, coords = [], []
    for line in lines[2:2 + n]:
        sym, x, y, z = line.split()[:4]
        atomnos.append(ATOMIC_NUMBERS[sym])
        coords.append([float(x), float(y), float(z)])
    return np.array(atomnos, dtype=np.uint64), np.array(coords, dtype=np.float64)


def write_xyz(path, atomnos, coords, comment=""):
    sym = {v: k for k, v in ATOMIC_NUMBERS.items()}
    with open(path, "w") as fh:
        fh.write(f"{len(atomnos)}\n{comment}\n")
        for z, (x, y, w) in zip(atomnos, coords):
            fh.write(f"{sym[int(z)]} {x:.15f} {y:.15f} {w:.15f}\n")


def molecule(spec):
    return (np.array([z for z, _ in spec], dtype=np.uint64),
            np.array([xyz for _, xyz in spec], dtype=np.float64))


def water_cluster(n_waters, seed=20261017, spacing=5.67):
    """(H2O)_n: monomer = water_crawford atoms; molecule i sits at lattice point
    (i mod g, (i//g) mod g, i//g^2)*spacing with g = ceil(n^(1/3)), rotated by a random unit
    quaternion drawn from numpy default_rng(seed)."""
    rng = np.random.default_rng(seed)
    g = max(1, math.ceil(round(n_waters ** (1.0 / 3.0), 9)))
    z0, x0 = molecule(WATER_CRAWFORD)
    atomnos, coords = [], []
    for i in range(n_waters):
        q = rng.standard_normal(4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        shift = np.array([i % g, (i // g) % g, i // (g * g)], dtype=np.float64) * spacing
        atomnos.extend(z0.tolist())
        coords.extend((x0 @ R.T + shift).tolist())
    return np.array(atomnos, dtype=np.uint64), np.array(coords, dtype=np.float64)


def synthetic_density(n, seed=20261017):
    """Symmetric test density D = (A + A^T)/2/n, A ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    return (A + A.T) / 2.0 / n
