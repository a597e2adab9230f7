"""Workloads of the per-ion deterministic criterion (shared by the CPU twin and the GPU test)."""
import os

import numpy as np

from mytrim_b200 import capi
from tests import util

# (name, primaries): sized so that the single-threaded FP32 host replay (tests/hostsim.cpp) takes ~1-2 s each
PER_ION_CASES = [("cu_on_cu_10keV", 800), ("cu_on_cu_1keV", 4000), ("h_on_fe_100keV", 1500), ("he_on_fe_100keV", 300),
                 ("c_on_w_1MeV", 48), ("xe_on_zro2_500keV", 12), ("cu_on_cu_150keV", 40), ("h_on_fe_1MeV", 300),
                 ("xe_on_uo2_10MeV", 2), ("uo2_fission_like", 160)]


def fission_like_primaries(n, seed=3):
    """Heterogeneous primaries: every ion has its own (Z, m), like mytrim_uo2's fission fragments
    (apps/mytrim_uo2.C:226-270), some starting inside a bubble."""
    rng = np.random.default_rng(seed)
    ions = capi.make_ions(n, 1, 1.0, 1.0)
    ions["Z"] = rng.integers(30, 62, n)
    ions["m"] = np.round(ions["Z"] * 2.55 + rng.uniform(-3, 3, n), 3)
    ions["E"] = rng.uniform(2e4, 2e5, n)
    ions["pos"] = rng.uniform(0, 400, (n, 3))
    d = rng.normal(size=(n, 3))
    ions["dir"] = d / np.linalg.norm(d, axis=1)[:, None]
    return ions


def setup_case(eng, name, n):
    """Materials + geometry of a case on any engine (CUDA, oracle, host replay); returns the primaries."""
    if name == "uo2_fission_like":
        cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
        eng.set_materials([util.UO2, util.XE_GAS])
        eng.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        ions = fission_like_primaries(n)
        k = min(n // 4, 40)
        ions["pos"][:k] = cl[np.arange(k) % len(cl), :3] + 2.0
        return ions
    c = util.setup_engine(eng, name)
    return util.primaries_for(c, n)
