"""Helpers to compare reach-set tables (neutral layout of armour_export_reachsets) with the fixtures under
tests/golden/ref/, which were generated from the REFERENCE's own sources by tools/make_golden.py."""
import glob
import os

import numpy as np

from conftest import GOLDEN

NF = 7
FIXTURES = sorted(glob.glob(os.path.join(GOLDEN, "ref", "*.npz")))


def load(path):
    return dict(np.load(path))


def check_tables(tb, gold, NJ=7, exact=True, coef_tol=0.0, rel_radius=0.0):
    """tb: full tables; gold: fixture.  exact=True demands bit-identical coefficients/centres (oracle);
    otherwise coefficients within coef_tol and radii >= the reference's and within rel_radius (GPU)."""
    ts = gold["t_subset"]
    sel_l = np.array([t * NJ + l for t in ts for l in range(NJ)])
    sel_u = np.array([t * NF + j for t in ts for j in range(NF)])
    assert np.array_equal(tb["nl"][sel_l], gold["nl"]), "link monomial counts"
    assert np.array_equal(tb["nu"][sel_u], gold["nu"]), "torque monomial counts"
    hl = np.concatenate([tb["hl"][i, :tb["nl"][i]] for i in sel_l]).astype(np.uint16)
    gl = np.concatenate([tb["gl"][i, :tb["nl"][i]] for i in sel_l])
    hu = np.concatenate([tb["hu"][i, :tb["nu"][i]] for i in sel_u]).astype(np.uint16)
    gu = np.concatenate([tb["gu"][i, :tb["nu"][i]] for i in sel_u])
    assert np.array_equal(hl, gold["hl"]) and np.array_equal(hu, gold["hu"]), "monomial key sets"

    def close(a, b, what):
        if exact:
            assert np.array_equal(a, b), what
        else:
            assert np.max(np.abs(a - b), initial=0.0) <= coef_tol, what

    close(gl, gold["gl"], "link coefficients")
    close(gu, gold["gu"], "torque coefficients")
    close(tb["cl"][sel_l], gold["cl"], "link centres")
    close(tb["cu"][sel_u], gold["cu"], "torque centres")
    G, Gr = np.asarray(tb["link_gens"]).reshape(-1, NJ, 6, 3)[ts], gold["link_gens"].reshape(-1, NJ, 6, 3)
    close(G[:, :, :3], Gr[:, :, :3], "link generator columns")
    pairs = [(tb["ru"][sel_u], gold["ru"], "u radius"),
             (np.asarray(tb["torque_radius"])[:, ts], gold["torque_radius"], "torque radius"),
             (G[:, :, 3:], Gr[:, :, 3:], "link radius")]
    for a, b, what in pairs:
        if exact:
            assert np.array_equal(a, b), what
        else:
            assert np.all(a >= b), what + ": does not contain the reference interval"
            nz = b > 0
            assert np.max((a[nz] - b[nz]) / b[nz], initial=0.0) <= rel_radius, what
