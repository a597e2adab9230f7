"""Workloads of the per-ion deterministic criterion (shared by the CPU twin and the GPU test)."""
import os

import numpy as np

from mytrim_b200 import capi
from tests import util

# (name, primaries): sized so that the single-threaded FP32 host replay (tests/hostsim.cpp) takes ~1-2 s each
PER_ION_CASES = [("cu_on_cu_10keV", 800), ("cu_on_cu_1keV", 4000), ("h_on_fe_100keV", 1500), ("he_on_fe_100keV", 300),
                 ("c_on_w_1MeV", 48), ("xe_on_zro2_500keV", 12), ("cu_on_cu_150keV", 40), ("h_on_fe_1MeV", 300),
                 ("xe_on_uo2_10MeV", 2), ("uo2_fission_like", 160),
                 # options and geometries of the all-options kernel (TrimBase::_potential, SimconfType::setLengthScale,
                 # SampleWire with CUT boundaries, SampleBurriedWire, a stack of different materials)
                 ("cu_1keV_moliere", 3000), ("cu_1keV_ckr", 3000), ("cu_20keV_scale10", 400), ("wire", 600),
                 ("burried_wire", 600), ("layer_stack", 600)]

# engine options of a case (tally mask etc. are added by the tests)
CASE_OPTIONS = {"cu_1keV_moliere": dict(potential=capi.POT_MOLIERE), "cu_1keV_ckr": dict(potential=capi.POT_CKR),
                "cu_20keV_scale10": dict(length_scale=10.0)}


def case_options(name):
    return dict(CASE_OPTIONS.get(name, {}))


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
    if name in ("cu_1keV_moliere", "cu_1keV_ckr"):
        c = util.setup_engine(eng, "cu_on_cu_1keV")
        return util.primaries_for(c, n)
    if name == "cu_20keV_scale10":
        # positions and the sample in units of 10 A, per-element Edisp / Elbind, Ef = 5 eV (make_golden.OPTION_CASE)
        eng.set_materials([{"rho": 8.92, "elements": [{"Z": 29, "m": 63.546, "t": 1.0, "Edisp": 30.0, "Elbind": 2.0}]}])
        eng.set_layers([100.0], wy=10.0, wz=10.0)
        return capi.make_ions(n, 29, 63.546, 2.0e4, pos=(0.0, 5.0, 5.0), Ef=5.0)
    if name in ("wire", "burried_wire"):
        from tests.golden.make_golden import GEOMETRY_CASES
        _, box, mats, ion, start, _ = GEOMETRY_CASES[name]
        eng.set_materials(mats)
        if name == "wire":
            eng.set_geometry(capi.GEOM_WIRE, box, bc=(capi.BC_CUT, capi.BC_CUT, capi.BC_PBC))
        else:
            eng.set_geometry(capi.GEOM_BURIED_WIRE, box, bc=(capi.BC_INF, capi.BC_INF, capi.BC_INF))
        return capi.make_ions(n, ion[0], ion[1], ion[2], pos=start[:3], direction=start[3:])
    if name == "layer_stack":
        from tests.golden.make_golden import STACK_CASE
        util.setup_engine(eng, STACK_CASE)
        return util.primaries_for(STACK_CASE, n)
    c = util.setup_engine(eng, name)
    return util.primaries_for(c, n)
