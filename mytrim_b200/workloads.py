"""The BASELINE.json configurations as concrete inputs for bench.py, the tools and the tests:
materials, sample, primaries, tally classes and launch sizes.  Host-side set-up only; everything runs through the
C ABI (capi.Engine)."""
import os

import numpy as np

from . import capi

CU = {"rho": 8.92, "elements": [{"Z": 29, "m": 63.546, "t": 1.0}]}
FE = {"rho": 7.8658, "elements": [{"Z": 26, "m": 55.847, "t": 1.0}]}
W = {"rho": 19.35, "elements": [{"Z": 74, "m": 183.85, "t": 1.0}]}
ZRO2 = {"rho": 6.52, "elements": [{"Z": 40, "m": 90.0, "t": 1.0}, {"Z": 8, "m": 16.0, "t": 2.0}]}
UO2 = {"rho": 10.0, "elements": [{"Z": 92, "m": 235.0, "t": 1.0}, {"Z": 8, "m": 16.0, "t": 2.0}]}
XE_GAS = {"rho": 3.5, "elements": [{"Z": 54, "m": 132.0, "t": 1.0}]}

CONFIGS = {
    "cu_on_cu_10keV": dict(ion=(29, 63.546, 1.0e4), materials=[CU], thicknesses=[1000.0]),
    "cu_on_cu_1keV": dict(ion=(29, 63.546, 1.0e3), materials=[CU], thicknesses=[100000.0]),
    "h_on_fe_100keV": dict(ion=(1, 1.008, 1.0e5), materials=[FE], thicknesses=[100000.0]),
    "he_on_fe_100keV": dict(ion=(2, 4.003, 1.0e5), materials=[FE], thicknesses=[100000.0]),
    "c_on_w_1MeV": dict(ion=(6, 12.0, 1.0e6), materials=[W], thicknesses=[10000.0]),
    "xe_on_uo2_80MeV": dict(ion=(54, 132.0, 8.0e7), materials=[UO2], thicknesses=[1.0e7]),
    # the file-energy / long-cascade configurations of SURVEY.md §8d (tests/json/cu_on_cu.json,
    # validation/h_on_fe at 1 MeV, tests/json/xe_on_uo2.json)
    "cu_on_cu_150keV": dict(ion=(29, 63.546, 1.5e5), materials=[CU], thicknesses=[1000.0]),
    "h_on_fe_1MeV": dict(ion=(1, 1.008, 1.0e6), materials=[FE], thicknesses=[1.0e6]),
    "xe_on_uo2_10MeV": dict(ion=(54, 131.904, 1.0e7), thicknesses=[100000.0], materials=[
        {"rho": 10.97, "elements": [{"Z": 92, "m": 238.03, "t": 1.0}, {"Z": 8, "m": 15.999, "t": 2.0}]}]),
    "xe_on_zro2_500keV": dict(ion=(54, 131.0, 5.0e5), materials=[ZRO2] * 50, thicknesses=[10.0] * 50,
                              box=(500.0, 100.0, 100.0)),
}

UO2_SEED = 39172          # tests/uo2/test.sh
UO2_BOX = (400.0, 400.0, 400.0)
# bubbles of the gold run (r = 10 A, Cbf = 0.1: 4 bubbles; tests/uo2/gold/out.clcoor)
UO2_CLUSTERS_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                 "uo2_out.clcoor")

# bench.py: BASELINE.json config -> (workload, primaries per GPU per launch, tally mask, what it is)
BENCH_WORKLOADS = {
    "cu_on_cu_10keV": dict(primaries=1 << 23, tally=capi.TALLY_VAC_DEPTH,
                           desc="Cu->Cu 10 keV, 1000 A Cu layer, full cascades, TrimVacCount tallies (validation/cu_on_cu)"),
    "h_on_fe_100keV": dict(primaries=1 << 23, tally=capi.TALLY_VAC_DEPTH,
                           desc="H->Fe 100 keV, full cascades, TrimVacCount tallies (validation/h_on_fe)"),
    "he_on_fe_100keV": dict(primaries=1 << 21, tally=capi.TALLY_VAC_DEPTH,
                            desc="He->Fe 100 keV, full cascades, TrimVacCount tallies (validation/he_on_fe)"),
    "c_on_w_1MeV": dict(primaries=1 << 18, tally=capi.TALLY_VAC_ENERGY,
                        desc="C->W 1 MeV, full cascades, TrimVacEnergyCount tallies (validation/c_on_w/input.json)"),
    "xe_on_zro2_500keV": dict(primaries=1 << 16, tally=capi.TALLY_VAC_DEPTH,
                              desc="Xe->ZrO2 500 keV, 50 x 10 A layers (inputs/samplelayers_zro2_multilayer.in), full cascades"),
    # configuration 4 as the reference's own driver runs it (apps/mytrim_layers.C:72,116-117: TrimRecoils = primaries and
    # first-generation recoils, Kinchin-Pease estimate for the rest)
    "xe_on_zro2_500keV_trimrecoils": dict(primaries=1 << 20, tally=0, config="xe_on_zro2_500keV", ref_tally="recoils",
                                          engine=dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP),
                                          desc="Xe->ZrO2 500 keV, 50 x 10 A layers, TrimRecoils as in apps/mytrim_layers.C "
                                               "(recoils of generation < 2 followed, Kinchin-Pease for generation 2)"),
    # the file-energy / long-cascade variants of configurations 1 and 2
    "cu_on_cu_150keV": dict(primaries=1 << 19, tally=capi.TALLY_VAC_DEPTH,
                            desc="Cu->Cu 150 keV, full cascades, TrimVacCount tallies (tests/json/cu_on_cu.json)"),
    "h_on_fe_1MeV": dict(primaries=1 << 22, tally=capi.TALLY_VAC_DEPTH,
                         desc="H->Fe 1 MeV, full cascades, TrimVacCount tallies (validation/h_on_fe at 1 MeV)"),
    "xe_on_uo2_10MeV": dict(primaries=1 << 13, tally=capi.TALLY_VAC_DEPTH,
                            desc="Xe->UO2 10 MeV, full cascades, TrimVacCount tallies (tests/json/xe_on_uo2.json)"),
    "uo2_fission": dict(primaries=1 << 16, tally=capi.TALLY_IONLOG, ionlog_z=54,
                        desc="fission-fragment pairs in UO2 with Xe bubbles (tests/uo2), Xe ion log"),
}


def setup_engine(eng, cfgname_or_dict):
    c = CONFIGS[cfgname_or_dict] if isinstance(cfgname_or_dict, str) else cfgname_or_dict
    eng.set_materials(c["materials"])
    box = c.get("box")
    if box:
        eng.set_layers(c["thicknesses"], wy=box[1], wz=box[2], wx=box[0])
    else:
        eng.set_layers(c["thicknesses"])
    return c


def primaries_for(c, n, seeds=None):
    Z, m, E = c["ion"]
    box = c.get("box")
    wy, wz = (box[1], box[2]) if box else (100.0, 100.0)
    return capi.make_ions(n, Z, m, E, pos=(0.0, wy / 2.0, wz / 2.0), seeds=seeds)


def setup_uo2(eng):
    """sampleClusters(400, 400, 400) with the gold run's bubbles: UO2 matrix, Xe gas (mytrim_uo2.C:124-184)."""
    cl = np.loadtxt(UO2_CLUSTERS_FILE)[:, :4]
    eng.set_materials([UO2, XE_GAS])
    eng.set_geometry(capi.GEOM_CLUSTERS, UO2_BOX, kn=(39, 39, 39), clusters=cl)


def setup_workload(eng, name, n, first_primary=0):
    """Sample + primaries [first_primary, first_primary + n) of a bench workload.  The beams are n identical
    primaries; the fission source draws event range [first/2, (first+n)/2) of its sequential stream."""
    if name == "uo2_fission":
        setup_uo2(eng)
        assert n % 2 == 0 and first_primary % 2 == 0
        return capi.fission_pairs(UO2_SEED, first_primary // 2, n // 2, UO2_BOX)
    c = setup_engine(eng, BENCH_WORKLOADS.get(name, {}).get("config", name))
    return primaries_for(c, n)
