"""Basis Set Exchange JSON ingestion: the on-disk format on the input side of
the path.  The reference deserialises `bse.get_basis(name, elements=[...], fmt="json")` into
BSEResult / BSEElement / BSEElectronShell (basis.rs:14-85) and then builds one CGTO per
Cartesian function in Basis::new (basis.rs:182-211).  This module does the same from a JSON
string, file or dict -- without needing the `basis_set_exchange` package at run time -- and
hands the flat CGTO arrays to the C ABI (`rchem_basis_create`).

Differences from the reference, on purpose:
  * `coefficients` is indexed by the POSITION of the angular momentum inside the shell, not by
    its value (basis.rs:193,201 index by value, which only works for s and sp shells), so lone
    p/d shells and general contractions load correctly; for s / sp shells both agree.
  * a general contraction (several coefficient rows for ONE angular momentum) becomes one CGTO
    set per row, in row order.
  * spherical shells ("gto_spherical") are rejected: the path is Cartesian only.
"""
import json

import numpy as np

from . import Basis, get_ijk_list, normalization


def _as_dict(source):
    if isinstance(source, dict):
        return source
    if isinstance(source, (bytes, bytearray)):
        source = source.decode()
    if isinstance(source, str) and source.lstrip().startswith("{"):
        return json.loads(source)
    with open(source) as fh:
        return json.load(fh)


def cgtos_from_bse(source, atomnos, all_atomcoords):
    """Flat CGTO arrays (origins, powers, prim_offset, exps, coefs, norms) in the reference's
    function order: atom -> electron shell -> angular momentum -> get_ijk_list component."""
    data = _as_dict(source)
    elements = data["elements"]
    coords = np.asarray(all_atomcoords, dtype=np.float64).reshape(-1, 3)
    origins, powers, off, exps, coefs, norms = [], [], [0], [], [], []
    for z, xyz in zip(atomnos, coords):
        entry = elements.get(str(int(z)))
        if entry is None:
            raise KeyError(f"element Z={int(z)} is not in the basis-set JSON")
        for shell in entry["electron_shells"]:
            if shell.get("function_type", "gto") == "gto_spherical":
                raise ValueError("spherical shells are not supported (Cartesian path)")
            shell_exps = [float(e) for e in shell["exponents"]]       # stored as strings
            rows = [[float(c) for c in row] for row in shell["coefficients"]]
            ams = [int(a) for a in shell["angular_momentum"]]
            if len(ams) == 1 and len(rows) > 1:                        # general contraction
                pairs = [(ams[0], row) for row in rows]
            elif len(ams) == len(rows):
                pairs = list(zip(ams, rows))
            else:
                raise ValueError("angular_momentum and coefficients do not line up")
            for am, row in pairs:
                if len(row) != len(shell_exps):
                    raise ValueError("one coefficient per exponent expected")  # basis.rs:191-194
                for pw in get_ijk_list(am):
                    origins.append(xyz)
                    powers.append(pw)
                    for e, c in zip(shell_exps, row):
                        exps.append(e)
                        coefs.append(c)
                        norms.append(normalization(pw, e))
                    off.append(len(exps))
    return (np.array(origins), np.array(powers, dtype=np.int32), np.array(off, dtype=np.int32),
            np.array(exps), np.array(coefs), np.array(norms))


def basis_from_bse(source, atomnos, all_atomcoords):
    """Basis::new with the basis set read from BSE JSON instead of the embedded tables."""
    return Basis.from_cgtos(*cgtos_from_bse(source, atomnos, all_atomcoords))
