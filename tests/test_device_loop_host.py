"""The CUDA kernel's lane loop (mytrim_b200/csrc/mtb_transport.cuh) compiled for the host
(tests/hostsim.cpp) against the double-precision oracle with the same Philox streams.  This checks
the FP32 reformulations, the depth-first stack traversal and the tallies without a GPU; the GPU
tests repeat it through the real kernels."""
import ctypes as C
import os

import numpy as np
import pytest

from mytrim_b200 import capi
from tests import util

TOL = 1e-5  # north_star: per-ion trajectories within 1e-5 relative


def _pair(cfg, name):
    orc = util.OracleEngine(util.ORC_RNG_PHILOX, **cfg)
    hs = util.HostSimEngine(**cfg)
    c = util.setup_engine(orc, name)
    util.setup_engine(hs, name)
    return orc, hs, c


@pytest.mark.parametrize("name,n", [("cu_on_cu_10keV", 200), ("cu_on_cu_1keV", 500), ("h_on_fe_100keV", 200),
                                    ("he_on_fe_100keV", 60), ("c_on_w_1MeV", 12), ("xe_on_zro2_500keV", 6)])
def test_trajectories_match_oracle(name, n):
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS | capi.TALLY_PHONON)
    orc, hs, c = _pair(cfg, name)
    ions = util.primaries_for(c, n)
    ro = orc.run(ions, seed=2344, records=True)
    rh = hs.run(ions, seed=2344, records=True)
    same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
    assert same.mean() >= 0.8
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    sel = ro["primary_steps"] == rh["primary_steps"]
    assert sel.mean() >= 0.95
    rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= max(2, 0.005 * len(rel)) and np.median(rel) < 0.1 * TOL
    bad = np.abs(ro["Eel"][same] - rh["Eel"][same]) > TOL * ro["Eel"][same]
    assert bad.sum() <= max(2, 0.005 * same.sum()), bad.sum()
    co, ch = orc.counters(), hs.counters()
    # energy partition closes: E0 = Eel + Enuc for every non-lost cascade
    E0 = c["ion"][2] * n
    assert abs(ch["EelTotal"] + ch["EnucTotal"] - E0) < 1e-6 * E0
    assert abs(co["EelTotal"] + co["EnucTotal"] - E0) < 1e-9 * E0
    assert ch["stack_max"] <= 32


def test_edge_cases_empty_single_and_degenerate_primaries():
    """The device loop on an empty batch, one primary, a primary without energy (parked as an interstitial
    without a collision), starts in front of / behind the layer stack, a sub-threshold primary and a
    direction that is not normalised."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    orc, hs, c = _pair(cfg, "cu_on_cu_10keV")
    rec = hs.run(util.primaries_for(c, 0), seed=3, records=True)
    assert len(rec) == 0 and hs.counters()["steps"] == 0
    one = util.primaries_for(c, 1)
    r1, o1 = hs.run(one, seed=3, records=True), orc.run(one, seed=3, records=True)
    assert r1["steps"][0] == o1["steps"][0] and r1["vacancies"][0] == o1["vacancies"][0]
    ions = util.primaries_for(c, 5)
    ions["E"][0] = 0.0
    ions["pos"][1] = (-40.0, 50.0, 50.0)
    ions["pos"][2] = (5000.0, 50.0, 50.0)
    ions["E"][3] = 10.0
    ions["dir"][4] = (0.0, 0.0, 2.0)
    r = hs.run(ions, seed=5, records=True)
    o = orc.run(ions[1:], seed=5, first_index=1, records=True)
    assert r["steps"][0] == 0 and r["state"][0] == capi.INTERSTITIAL and r["E"][0] == 0.0
    assert np.array_equal(r["steps"][1:], o["steps"]) and np.array_equal(r["vacancies"][1:], o["vacancies"])
    assert np.array_equal(r["state"][1:], o["state"])
    d = np.abs(r["pos"][1:] - o["pos"]).max(axis=1)
    assert (d < 1e-5 * np.maximum(np.abs(o["pos"]).max(axis=1), 1.0)).all()
    assert r["vacancies"][3] == 0 and r["ions"][3] == 1


def test_stopping_matches_oracle():
    import json, os
    data = json.load(open(os.path.join(util.GOLDEN, "stopping.json")))
    for name, d in data.items():
        E = np.logspace(1.5, 8.5, 200)
        with util.OracleEngine(util.ORC_RNG_PHILOX) as orc, util.HostSimEngine() as hs:
            orc.set_materials([d["material"]])
            hs.set_materials([d["material"]])
            a = orc.stopping(0, d["Z"], d["m"], E)
            b = hs.stopping(0, d["Z"], d["m"], E)
        assert np.abs(b / a - 1).max() < 5e-6, name


def test_follow_policies_and_vacancy_models():
    for cfg in (dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT, tally_mask=capi.TALLY_RANGE),
                dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP),
                dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=1, vacancy_model=capi.VAC_KP),
                dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VACMAP, vmap_z=(29, 8, -1))):
        orc, hs, c = _pair(cfg, "cu_on_cu_10keV")
        ions = util.primaries_for(c, 150)
        orc.run(ions, seed=11)
        hs.run(ions, seed=11)
        co, ch = orc.counters(), hs.counters()
        for k in ("steps", "ions", "replacements", "recoils_queued"):
            assert abs(co[k] - ch[k]) <= 2e-3 * co[k], (cfg, k)   # a flipped branch changes a cascade
        assert abs(co["vacancies_created"] - ch["vacancies_created"]) <= 2e-4 * co["vacancies_created"] + 1
        if cfg.get("tally_mask", 0) & capi.TALLY_RANGE:
            xo, zo = orc.range_list()
            xh, zh = hs.range_list()
            assert abs(len(xo) - len(xh)) <= 2e-3 * len(xo)
            assert abs(np.mean(xo) - np.mean(xh)) < 1e-2 * abs(np.mean(xo))
        if cfg.get("tally_mask", 0) & capi.TALLY_VAC_ENERGY:
            eo, eh = orc.vac_energy(), hs.vac_energy()
            assert abs(int(eo.sum()) - int(eh.sum())) <= 2e-3 * eo.sum()
            assert np.abs(eo.astype(int) - eh.astype(int)).sum() <= 2e-2 * eo.sum()
            assert np.abs(orc.vacmap().astype(int) - hs.vacmap().astype(int)).sum() <= 2e-2 * orc.vacmap().sum()


def test_ion_log_and_events():
    cfg = dict(tally_mask=capi.TALLY_IONLOG, ionlog_z=8)
    orc, hs, c = _pair(cfg, "xe_on_zro2_500keV")
    ions = util.primaries_for(c, 2)
    orc.run(ions, seed=5)
    hs.run(ions, seed=5)
    lo, lh = orc.ion_log(), hs.ion_log()
    assert len(lo) == len(lh) > 100
    lo, lh = np.sort(lo, order="uid"), np.sort(lh, order="uid")
    assert np.array_equal(lo["uid"], lh["uid"]) and np.array_equal(lo["gen"], lh["gen"])
    assert (lo["Z"] == 8).all() and np.array_equal(lo["state"], lh["state"])
    # a rare flipped branch (Newton iteration count, threshold test) moves single ions by more
    d = np.abs(lo["pos1"] - lh["pos1"]).max(axis=1)
    assert np.quantile(d, 0.995) < 1e-3 and d.max() < 1.0
    # single-ion event mode (mtb_trim_one): the hooks' view of every collision
    ion = ions[0]
    fo, so, eo = orc.trim_one(ion, 99, 1234)
    fh, sh, eh = hs.trim_one(ion, 99, 1234)
    assert so == sh and len(eo) == len(eh) > 10
    for f in ("material", "element", "pka_state", "recoil_above_threshold"):
        assert np.array_equal(eo[f], eh[f]), f
    assert np.abs(eo["pka_pos"] - eh["pka_pos"]).max() < 1e-5 * np.abs(eo["pka_pos"]).max()
    assert np.abs(eo["recoil_E"] - eh["recoil_E"]).max() <= 1e-5 * np.abs(eo["recoil_E"]).max()
    assert np.abs(fo["pos"] - fh["pos"]).max() < 1e-5 * np.abs(fo["pos"]).max()


def _fission_like_primaries(n, seed=3):
    """Heterogeneous primaries: every ion has its own (Z, m), like mytrim_uo2's fission fragments."""
    rng = np.random.default_rng(seed)
    ions = capi.make_ions(n, 1, 1.0, 1.0)
    ions["Z"] = rng.integers(30, 62, n)
    ions["m"] = np.round(ions["Z"] * 2.55 + rng.uniform(-3, 3, n), 3)
    ions["E"] = rng.uniform(2e4, 2e5, n)
    ions["pos"] = rng.uniform(0, 400, (n, 3))
    d = rng.normal(size=(n, 3))
    ions["dir"] = d / np.linalg.norm(d, axis=1)[:, None]
    return ions


def test_per_primary_species_and_clusters():
    """More distinct primary species than the class table holds: the lane builds private rows.
    Clusters geometry (tests/uo2: UO2 matrix, Xe bubbles) at the same time."""
    cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_PHONON)
    ions = _fission_like_primaries(48)
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, util.HostSimEngine(**cfg) as hs:
        for e in (orc, hs):
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        ro = orc.run(ions, seed=17, records=True)
        rh = hs.run(ions, seed=17, records=True)
        co, ch = orc.counters(), hs.counters()
    same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
    assert same.mean() >= 0.8, same.mean()
    sel = ro["primary_steps"] == rh["primary_steps"]
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= 2 and np.median(rel) < 0.1 * TOL
    E0 = ions["E"].sum()
    assert abs(ch["EelTotal"] + ch["EnucTotal"] - E0) < 1e-6 * E0
    assert abs(ch["steps"] - co["steps"]) <= 0.02 * co["steps"]


@pytest.mark.parametrize("bc,ncl,kn,radii", [((capi.BC_PBC,) * 3, 300, 20, (6.0, 19.0)),
                                              ((capi.BC_CUT, capi.BC_PBC, capi.BC_INF), 300, 20, (6.0, 19.0)),
                                              ((capi.BC_PBC,) * 3, 80, 39, (9.0, 10.0))])
def test_dense_clusters_neighbourhood_filter(bc, ncl, kn, radii):
    """Bubbles dense enough that ions meet them.  300 in a 20^3 hash: scan neighbourhoods overlap, chains
    form, the box faces matter — the distance map in front of the 27-cell scan must not change a single
    lookup.  80 in a 39^3 hash: cells up to ~6 cells from the nearest bubble, so lanes skip lookups for
    tens of Angstrom of path (cl_dist) — and must look again before a bubble can be reached."""
    rng = np.random.default_rng(11)
    cl = np.column_stack([rng.uniform(0, 400, (ncl, 3)), rng.uniform(radii[0], radii[1], ncl)])
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_PHONON | capi.TALLY_IONLOG, ionlog_z=54)
    n = 160
    ions = capi.make_ions(n, 36, 84.0, 3.0e5)   # Kr: every Xe ion in the log is a displaced bubble atom
    ions["pos"] = rng.uniform(0, 400, (n, 3))
    ions["pos"][:20] = cl[:20, :3] + rng.uniform(-3, 3, (20, 3))   # some start inside a bubble
    d = rng.normal(size=(n, 3))
    ions["dir"] = d / np.linalg.norm(d, axis=1)[:, None]
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, util.HostSimEngine(**cfg) as hs:
        for e in (orc, hs):
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), bc=bc, kn=(kn, kn, kn), clusters=cl)
        ro = orc.run(ions, seed=5, records=True)
        rh = hs.run(ions, seed=5, records=True)
        co, ch = orc.counters(), hs.counters()
        lo, lh = orc.ion_log(), hs.ion_log()
    same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
    assert same.mean() >= 0.8, same.mean()
    assert (ro["state"] != rh["state"]).sum() <= 2
    sel = ro["primary_steps"] == rh["primary_steps"]
    assert sel.mean() >= 0.95
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= 2 and np.median(rel) < 0.1 * TOL
    assert abs(ch["steps"] - co["steps"]) <= 0.02 * co["steps"]
    assert abs(ch["left_sample"] - co["left_sample"]) <= 2 and abs(ch["lost"] - co["lost"]) <= 2
    # bubble atoms set in motion, per bubble of origin (the tag of a recoil born in a bubble is its index):
    # a lookup skipped or filtered by mistake would lose them
    assert len(lo) > 500, len(lo)
    ho = np.bincount(lo["tag"][lo["tag"] >= 0], minlength=ncl)
    hh = np.bincount(lh["tag"][lh["tag"] >= 0], minlength=ncl)
    assert np.abs(ho - hh).sum() <= 0.02 * ho.sum() + 2, (ho.sum(), hh.sum(), np.abs(ho - hh).sum())


@pytest.mark.parametrize("phonon", [True, False])
def test_clusters_variant_equals_generic(phonon):
    """The lean clusters variants of the lane loop (sampleClusters geometry, per-primary species, ion log /
    energy partition compiled in; everything else compiled out) against the all-options loop: identical.
    phonon=False is the mask of the tests/uo2 driver itself (ion log only): the CLUSTERS-LOG variant, tallies
    fixed at compile time."""
    cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
    cfg = dict(tally_mask=capi.TALLY_RECORDS | (capi.TALLY_PHONON if phonon else 0) | capi.TALLY_IONLOG, ionlog_z=54)
    ions = _fission_like_primaries(40)
    ions["pos"][:8] = cl[np.arange(8) % len(cl), :3] + 2.0   # some start inside a bubble
    with util.HostSimEngine(**cfg) as a, util.HostSimEngine(**cfg) as b:
        for e in (a, b):
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        b._lib.hs_force_generic(b._h, 1)
        ra = a.run(ions, seed=8, records=True)
        rb = b.run(ions, seed=8, records=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), f
        ca, cb = a.counters(), b.counters()
        for k in ca:
            if k != "stack_max":   # diagnostic of the all-options loop only
                assert abs(ca[k] - cb[k]) <= 1e-12 * abs(cb[k]), k
        la, lb = a.ion_log(), b.ion_log()
        assert len(la) == len(lb) > 0
        for f in la.dtype.names:
            assert np.array_equal(np.sort(la[f], axis=0), np.sort(lb[f], axis=0)), f


@pytest.mark.parametrize("cfg", [
    dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT, tally_mask=capi.TALLY_RANGE | capi.TALLY_RECORDS),
    dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP, tally_mask=capi.TALLY_RECORDS),
    dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VAC_DEPTH | capi.TALLY_VACMAP | capi.TALLY_RECORDS),
])
def test_layers_variant_equals_generic(cfg):
    """The layered-sample variant (run-time follow policy, vacancy model and tallies; no clusters, wires,
    CUT boundaries, other potentials or per-primary species compiled in) against the all-options loop."""
    with util.HostSimEngine(**cfg) as a, util.HostSimEngine(**cfg) as b:
        for e in (a, b):
            c = util.setup_engine(e, "xe_on_zro2_500keV")
        b._lib.hs_force_generic(b._h, 1)
        ions = util.primaries_for(c, 4)
        ions["E"] = 6.0e4
        ra = a.run(ions, seed=12, records=True)
        rb = b.run(ions, seed=12, records=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), f
        ca, cb = a.counters(), b.counters()
        for k in ca:
            if k != "stack_max":
                assert abs(ca[k] - cb[k]) <= 1e-12 * abs(cb[k]), k
        if cfg["tally_mask"] & capi.TALLY_VAC_DEPTH:
            assert np.array_equal(a.vac_depth()[0], b.vac_depth()[0])
            assert np.array_equal(a.vac_energy(), b.vac_energy())
        if cfg["tally_mask"] & capi.TALLY_RANGE:
            xa, za = a.range_list()
            xb, zb = b.range_list()
            assert np.array_equal(np.sort(xa), np.sort(xb)) and len(xa) > 0


def test_mono_evac_variant_equals_generic():
    """Single-element sample + TrimVacEnergyCount tally (validation/c_on_w/input.json) selects the MONO-EVAC
    variant: identical to the all-options loop, 2-D tally included."""
    cfg = dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_RECORDS)
    with util.HostSimEngine(**cfg) as a, util.HostSimEngine(**cfg) as b:
        for e in (a, b):
            c = util.setup_engine(e, "c_on_w_1MeV")
        b._lib.hs_force_generic(b._h, 1)
        ions = util.primaries_for(c, 6)
        ions["E"] = 1.0e5
        ra = a.run(ions, seed=21, records=True)
        rb = b.run(ions, seed=21, records=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), f
        ca, cb = a.counters(), b.counters()
        for k in ca:
            if k != "stack_max":
                assert abs(ca[k] - cb[k]) <= 1e-12 * abs(cb[k]), k
        ea, eb = a.vac_energy(), b.vac_energy()
        assert np.array_equal(ea, eb) and ea.sum() > 0


def test_fast_phonon_variant_equals_generic():
    """Layered compound sample + the energy partition of TrimPhononOut selects the FAST-PHONON variant: identical
    to the all-options loop, and the partition closes (Eel + Enuc = energy of the primaries)."""
    cfg = dict(tally_mask=capi.TALLY_PHONON | capi.TALLY_RECORDS)
    with util.HostSimEngine(**cfg) as a, util.HostSimEngine(**cfg) as b:
        for e in (a, b):
            c = util.setup_engine(e, "xe_on_zro2_500keV")
        b._lib.hs_force_generic(b._h, 1)
        ions = util.primaries_for(c, 4)
        ions["E"] = 6.0e4
        ra = a.run(ions, seed=13, records=True)
        rb = b.run(ions, seed=13, records=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), f
        ca, cb = a.counters(), b.counters()
        for k in ca:
            if k != "stack_max":
                assert abs(ca[k] - cb[k]) <= 1e-12 * abs(cb[k]), k
        E0 = ions["E"].sum()
        assert ca["EnucTotal"] > 0 and abs(ca["EelTotal"] + ca["EnucTotal"] - E0) < 1e-6 * E0


def test_fast_kernel_defers_unknown_species():
    """Layered sample + TrimVacCount tallies selects the compile-time fast loop; primaries whose
    species has no class are handed to the generic loop and the union equals a generic-only run."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    ions = _fission_like_primaries(60)
    ions["pos"] = (0.0, 50.0, 50.0)
    ions["dir"] = (1.0, 0.0, 0.0)
    ions[::3] = capi.make_ions(20, 29, 63.546, 1.0e4)   # every third primary is a plain Cu ion
    with util.HostSimEngine(**cfg) as a, util.HostSimEngine(**cfg) as b:
        for e in (a, b):
            util.setup_engine(e, "cu_on_cu_10keV")
        b._lib.hs_force_generic(b._h, 1)
        ra = a.run(ions, seed=4, records=True)
        rb = b.run(ions, seed=4, records=True)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]), f
        ca, cb = a.counters(), b.counters()
        for k in ca:
            if k not in ("stack_max",):
                assert abs(ca[k] - cb[k]) <= 1e-12 * abs(cb[k]), k   # f64 totals: summation order differs
        assert np.array_equal(a.vac_depth()[0], b.vac_depth()[0])


@pytest.mark.parametrize("opts", [dict(potential=capi.POT_MOLIERE), dict(potential=capi.POT_CKR),
                                  dict(length_scale=10.0), dict(tmin=1.0, cw=0.01)])
def test_options_potentials_and_scale(opts):
    """MOLIERE / C-Kr potentials (trim.C:206-222, 247-259), SimconfType::setLengthScale, tmin/cw."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS, **opts)
    mat = {"rho": 8.92, "elements": [{"Z": 29, "m": 63.546, "t": 1.0, "Edisp": 30.0, "Elbind": 2.0}]}
    scale = opts.get("length_scale", 1.0)
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, util.HostSimEngine(**cfg) as hs:
        for e in (orc, hs):
            e.set_materials([mat])
            e.set_layers([1000.0 / scale], wy=100.0 / scale, wz=100.0 / scale)
        ions = capi.make_ions(150, 29, 63.546, 2.0e4, pos=(0.0, 50.0 / scale, 50.0 / scale), Ef=5.0)
        ro = orc.run(ions, seed=21, records=True)
        rh = hs.run(ions, seed=21, records=True)
        co, ch = orc.counters(), hs.counters()
    same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
    assert same.mean() >= 0.9, same.mean()
    sel = ro["primary_steps"] == rh["primary_steps"]
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0 / scale)
    rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= 2 and np.median(rel) < 0.1 * TOL
    assert abs(co["vacancies_created"] - ch["vacancies_created"]) <= 0.01 * co["vacancies_created"]
    if "length_scale" in opts:
        # positions are in units of 10 A: the same physics lands in 10x fewer depth bins
        assert ro["pos"][:, 0].mean() < 20.0


def test_mono_variant_equals_fast_variant_on_the_host():
    """A single-element sample takes the MONO variant (no geometry look-up, vacuum test, target pick, element
    loop); MYTRIM_B200_NO_MONO routes it through FAST.  Same arithmetic: without FMA contraction (host build)
    the per-primary records and the tallies are identical bit for bit."""
    import os
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    out = []
    for knob in (None, "MYTRIM_B200_NO_MONO"):
        if knob:
            os.environ[knob] = "1"
        try:
            res = {}
            for name, n in (("cu_on_cu_10keV", 150), ("h_on_fe_100keV", 100), ("c_on_w_1MeV", 6)):
                with util.HostSimEngine(**cfg) as hs:
                    c = util.setup_engine(hs, name)
                    r = hs.run(util.primaries_for(c, n), seed=77, records=True)
                    res[name] = (r.copy(), hs.vac_depth()[0].copy(), hs.counters())
            out.append(res)
        finally:
            if knob:
                os.environ.pop(knob)
    for name in out[0]:
        (ra, va, ca), (rb, vb, cb) = out[0][name], out[1][name]
        assert ra.tobytes() == rb.tobytes(), name
        assert np.array_equal(va, vb) and ca["steps"] == cb["steps"] and ca["vacancies_created"] == cb["vacancies_created"]


def test_wire_and_buried_wire_geometry_on_the_host():
    """SampleWire (CUT boundaries, vacuum outside the cylinder) and SampleBurriedWire (INF boundaries, cover layer,
    matrix) through the device loop (generic variant: F_GEOM_ANY, F_CUT) against the oracle on the same Philox
    streams; the oracle's geometry is pinned against the reference in test_oracle_golden.py."""
    cfg = dict(tally_mask=capi.TALLY_RECORDS)
    cases = [
        (capi.GEOM_WIRE, (60.0, 60.0, 1000.0), (capi.BC_CUT, capi.BC_CUT, capi.BC_PBC), [util.CU],
         dict(pos=(30.0, 18.0, 0.0), direction=(0.0, 0.3, 1.0))),
        (capi.GEOM_BURIED_WIRE, (100.0, 100.0, 300.0), (capi.BC_INF,) * 3, [util.CU, util.FE],
         dict(pos=(50.0, 50.0, -200.0), direction=(0.1, 0.0, 1.0))),
    ]
    for kind, box, bc, mats, start in cases:
        with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, util.HostSimEngine(**cfg) as hs:
            for e in (orc, hs):
                e.set_materials(mats)
                e.set_geometry(kind, box, bc=bc)
            ions = capi.make_ions(120, 29, 63.546, 2.0e4, **start)
            ro = orc.run(ions, seed=13, records=True)
            rh = hs.run(ions, seed=13, records=True)
            co, ch = orc.counters(), hs.counters()
        assert len(set(ro["state"].tolist())) >= 2
        assert (ro["state"] == rh["state"]).mean() > 0.97
        same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
        assert same.mean() >= 0.85, same.mean()
        for k in ("steps", "ions", "left_sample", "lost", "vacancies_created"):
            assert abs(ch[k] - co[k]) <= 0.02 * co[k] + 2, (kind, k, ch[k], co[k])
        sel = (ro["primary_steps"] == rh["primary_steps"]) & (ro["state"] == rh["state"])
        path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
        rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
        assert (rel >= TOL).sum() <= max(2, 0.01 * len(rel)) and np.median(rel) < 0.1 * TOL


@pytest.mark.parametrize("name,n", [("cu_on_cu_1keV", 4000), ("h_on_fe_100keV", 2500), ("cu_on_cu_10keV", 600)])
def test_statistics_of_the_host_loop_against_reference_golden(name, n):
    """CPU twin of the GPU statistical test: the FP32 device loop (host build, Philox streams, azimuth instead of
    the reference's rejection loop) against per-primary records of the UNMODIFIED reference (tests/golden):
    two-sample KS on projected range, lateral range, vacancies and electronic loss, means within 4 standard errors."""
    from scipy import stats
    gold = np.load(os.path.join(util.GOLDEN, "ref_records_%s.npz" % name))["records"]
    c = util.CONFIGS[name]
    with util.HostSimEngine(tally_mask=capi.TALLY_RECORDS) as hs:
        util.setup_engine(hs, c)
        rec = hs.run(util.primaries_for(c, n), seed=4242, records=True)
    for field, getter in (("x", lambda r: r["pos"][:, 0]), ("vacancies", lambda r: r["vacancies"].astype(float)),
                          ("Eel", lambda r: r["Eel"]),
                          ("lateral", lambda r: np.hypot(r["pos"][:, 1] - 50.0, r["pos"][:, 2] - 50.0))):
        a, b = getter(rec), getter(gold)
        p = stats.ks_2samp(a, b).pvalue
        assert p > 0.001, (name, field, p)
        se = np.sqrt(a.var() / len(a) + b.var() / len(b))
        assert abs(a.mean() - b.mean()) <= 4.0 * se + 1e-12, (name, field, a.mean(), b.mean(), se)


@pytest.mark.skipif(not util.have_reference(), reason="oracle/_ref (the compiled reference) is not built")
@pytest.mark.parametrize("name,n", [("cu_on_cu_1keV", 30000), ("h_on_fe_100keV", 12000)])
def test_north_star_statistical_criterion_on_the_host_loop(name, n):
    """The north-star statistical criterion (two-sample KS p > 0.01, means within 1 %) with the FP32 device loop
    built for the host against the UNMODIFIED reference run here (oracle/_ref, distinct 32-bit seeds, all cores):
    what a kernel change can be checked against when no GPU is at hand."""
    from scipy import stats
    c = util.CONFIGS[name]
    ref, _, _ = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], util.distinct_seeds(n, master=9),
                                            threads=os.cpu_count() or 1, box=c.get("box"))
    with util.HostSimEngine(tally_mask=capi.TALLY_RECORDS) as hs:
        util.setup_engine(hs, c)
        rec = hs.run(util.primaries_for(c, n), seed=777, records=True)
    for field, getter in (("x", lambda r: r["pos"][:, 0]), ("vacancies", lambda r: r["vacancies"].astype(float)),
                          ("replacements", lambda r: r["replacements"].astype(float)), ("Eel", lambda r: r["Eel"]),
                          ("steps", lambda r: r["steps"].astype(float)),
                          ("lateral", lambda r: np.hypot(r["pos"][:, 1] - 50.0, r["pos"][:, 2] - 50.0))):
        a, b = getter(rec), getter(ref)
        p = stats.ks_2samp(a, b).pvalue
        assert p > 0.01, (name, field, p)
        se = np.sqrt(a.var() / len(a) + b.var() / len(b))
        assert abs(a.mean() - b.mean()) <= max(0.01 * abs(b.mean()), 4.0 * se), (name, field, a.mean(), b.mean())


def test_stack_of_different_materials_on_the_host():
    """Layer look-up over DIFFERENT materials (binary search over the cumulative thicknesses; Cu / Fe / W / ZrO2) in
    the device loop against the oracle, whose look-up is pinned against the reference on the same stack."""
    from tests.golden.make_golden import STACK_CASE as o
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, util.HostSimEngine(**cfg) as hs:
        for e in (orc, hs):
            util.setup_engine(e, o)
        ions = util.primaries_for(o, 200)
        ro = orc.run(ions, seed=5, records=True)
        rh = hs.run(ions, seed=5, records=True)
        vo, vh = orc.vac_depth()[0], hs.vac_depth()[0]
    same = (ro["vacancies"] == rh["vacancies"]) & (ro["steps"] == rh["steps"]) & (ro["ions"] == rh["ions"])
    assert same.mean() >= 0.8, same.mean()
    sel = ro["primary_steps"] == rh["primary_steps"]
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    rel = (np.linalg.norm(ro["pos"] - rh["pos"], axis=1) / path)[sel]
    assert sel.mean() >= 0.95 and (rel >= TOL).sum() <= 2 and np.median(rel) < 0.1 * TOL
    n = max(len(vo), len(vh))
    assert np.abs(np.pad(vo, (0, n - len(vo))).astype(int) - np.pad(vh, (0, n - len(vh))).astype(int)).sum() <= 0.02 * vo.sum()


def test_host_loop_against_the_1e6_reference_summary():
    """The fixture of the GPU suite's 1e6-ion criterion (tests/golden/ref_stats_*.npz) with a smaller sample of
    the host build of the loop: KS p > 0.01, means within 1 % (4 standard errors for the noisier observables).
    The full-size run of the host loop is recorded in profiles/r01_statistics_host_loop_cu_on_cu_10keV.log."""
    summary = np.load(os.path.join(util.GOLDEN, "ref_stats_cu_on_cu_10keV.npz"))
    c = util.CONFIGS["cu_on_cu_10keV"]
    n = 8000
    with util.HostSimEngine(tally_mask=capi.TALLY_RECORDS) as hs:
        util.setup_engine(hs, c)
        rec = hs.run(util.primaries_for(c, n), seed=2344, records=True)
    for name, (mean, ref_mean, D, p) in util.ks_against_summary(rec, summary).items():
        var = float(summary["m_" + name][1])
        assert p > 0.01, (name, D, p)
        assert abs(mean - ref_mean) <= max(0.01 * abs(ref_mean), 4.0 * np.sqrt(var / n)), (name, mean, ref_mean)


@pytest.mark.parametrize("name,n", [(c, max(1, k // 4)) for c, k in __import__("tests.parity_cases", fromlist=["x"]).PER_ION_CASES])
def test_per_ion_parity_fp32_replay_against_fp64_oracle(name, n):
    """CPU twin of tests/test_gpu_parity.py::test_per_ion_parity_with_fp32_host_replay: EVERY followed ion of the FP32
    replay of the device loop against the FP64 oracle on the same Philox streams (ion log joined by ion id).
    Measured in this container (full case sizes, 1.0e6 ions in total): every ion has its twin with identical integer
    fields; 35 ions end more than 1e-5 (of the distance from the source) away from their twin."""
    from tests import parity_cases
    cfg = dict(tally_mask=capi.TALLY_IONLOG | capi.TALLY_RECORDS, ionlog_capacity=1 << 21, **parity_cases.case_options(name))
    with util.HostSimEngine(**cfg) as hs, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        ions = parity_cases.setup_case(hs, name, n)
        parity_cases.setup_case(orc, name, n)
        rh = hs.run(ions, seed=2344, records=True)
        ro = orc.run(ions, seed=2344, records=True)
        s = util.compare_ion_logs(hs.ion_log(1 << 21), orc.ion_log(1 << 21), ions)
        r = util.compare_records(rh, ro, ions)
    assert s["n_test"] == s["n_replay"] == s["joined"] > 0, s
    assert s["ints_equal"] >= 0.9999 * s["joined"], s
    assert s["pos_outliers"] <= 5e-4 * s["joined"] + 2 and s["energy_outliers"] <= 2, s
    assert s["median_rel_pos"] < 0.1 * TOL, s
    assert r["cascades_identical"] >= 0.9 * r["n"] and r["pos_outliers"] == 0, r


def test_identical_layer_materials_are_folded():
    """inputs/samplelayers_zro2_multilayer.in stacks 50 layers of one ZrO2: the host folds identical materials, the device
    sees ONE material (no layer search, class tables of one material); a stack of different materials keeps them, and a
    repeated material in such a stack maps onto its first occurrence (mtb_tables.h: fold_identical_materials)."""
    import ctypes as C
    from tests.golden.make_golden import STACK_CASE
    info = (C.c_int32 * 4)()
    with util.HostSimEngine(tally_mask=capi.TALLY_VAC_DEPTH) as hs:
        util.setup_engine(hs, "xe_on_zro2_500keV")
        assert hs._lib.hs_sample_info(hs._h, info) == 0
        assert list(info) == [1, 1, 0, 50]
    with util.HostSimEngine(tally_mask=capi.TALLY_VAC_DEPTH) as hs:
        util.setup_engine(hs, "cu_on_cu_10keV")
        assert hs._lib.hs_sample_info(hs._h, info) == 0 and list(info)[:3] == [1, 1, 1]
    with util.HostSimEngine(tally_mask=capi.TALLY_VAC_DEPTH) as hs:
        util.setup_engine(hs, STACK_CASE)
        assert hs._lib.hs_sample_info(hs._h, info) == 0 and list(info) == [4, 0, 0, 4]
    # A B A B: two device materials, and the same cascades as the unfolded oracle
    stack = dict(ion=(29, 63.546, 2.0e4), materials=[util.CU, util.FE, util.CU, util.FE], thicknesses=[30.0, 40.0, 30.0, 500.0])
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with util.HostSimEngine(**cfg) as hs, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        util.setup_engine(hs, stack)
        util.setup_engine(orc, stack)
        assert hs._lib.hs_sample_info(hs._h, info) == 0 and list(info) == [2, 0, 0, 4]
        ions = util.primaries_for(stack, 200)
        rh, ro = hs.run(ions, seed=4, records=True), orc.run(ions, seed=4, records=True)
        r = util.compare_records(rh, ro, ions)
        assert r["cascades_identical"] >= 198 and r["pos_outliers"] == 0, r
        # getrstop of input material 2 (= the folded Cu) and 3 (= Fe)
        E = np.array([1e3, 1e5])
        assert np.allclose(hs.stopping(2, 29, 63.546, E), orc.stopping(2, 29, 63.546, E), rtol=1e-5)
        assert np.allclose(hs.stopping(3, 29, 63.546, E), orc.stopping(3, 29, 63.546, E), rtol=1e-5)


def test_invalid_primaries_are_skipped_and_reported():
    """Host twin of the device-side validation: Z outside 1..92, m <= 0, NaN energy and a zero direction never reach the
    Z-indexed tables; the run reports MTB_EINVAL and the other primaries are followed."""
    bad = capi.make_ions(32, 29, 63.546, 1e3)
    bad["Z"][3] = 0
    bad["Z"][7] = 200
    bad["m"][11] = -1.0
    bad["E"][13] = np.nan
    bad["dir"][17] = 0.0
    with util.HostSimEngine(tally_mask=capi.TALLY_VAC_DEPTH) as hs:
        util.setup_engine(hs, "cu_on_cu_1keV")
        with pytest.raises(capi.MytrimError) as e:
            hs.run(bad, seed=1)
        assert e.value.code == capi.EINVAL
        assert hs.counters()["primaries"] == 27
        hs.run(capi.make_ions(8, 29, 63.546, 1e3), seed=1)
        assert hs.counters()["primaries"] == 35


def test_results_do_not_depend_on_what_the_engine_ran_before():
    """The first primaries of every batch register their species as projectile classes (table rows in shared memory
    instead of per-lane rows).  Which species are registered therefore depends on the handle's history — and must not
    change any result: mytrim_uo2 deals chunks of fission events over several GPUs and writes identical files."""
    from tests import parity_cases
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_PHONON)
    with util.HostSimEngine(**cfg) as used, util.HostSimEngine(**cfg) as fresh:
        first = parity_cases.setup_case(used, "uo2_fission_like", 40)
        parity_cases.setup_case(fresh, "uo2_fission_like", 40)
        second = parity_cases.fission_like_primaries(40, seed=11)
        used.run(first, seed=5, first_index=0)                       # registers the species of `first`
        ra = used.run(second, seed=5, first_index=1000, records=True)
        rb = fresh.run(second, seed=5, first_index=1000, records=True)   # registers the species of `second`
    for f in ra.dtype.names:
        assert np.array_equal(ra[f], rb[f]), f


VARIANT_CASES = [
    # (sample / workload, engine configuration, variant pick_variant() must select)
    ("cu_on_cu_10keV", dict(tally_mask=capi.TALLY_VAC_DEPTH), "MONO"),
    ("h_on_fe_100keV", dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS), "MONO"),
    ("c_on_w_1MeV", dict(tally_mask=capi.TALLY_VAC_ENERGY), "MONO-EVAC"),                      # validation/c_on_w/input.json
    ("xe_on_zro2_500keV", dict(tally_mask=capi.TALLY_VAC_DEPTH), "FAST"),                       # compound stack, folded
    ("xe_on_zro2_500keV", dict(tally_mask=capi.TALLY_PHONON | capi.TALLY_RECORDS), "FAST-PHONON"),
    ("xe_on_zro2_500keV", dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP), "LAYERS-PLAIN"),   # mytrim_layers
    ("xe_on_zro2_500keV", dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP,
                               tally_mask=capi.TALLY_RECORDS | capi.TALLY_VACMAP), "LAYERS"),
    ("cu_on_cu_10keV", dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT, tally_mask=capi.TALLY_RANGE), "LAYERS"),
    ("cu_on_cu_10keV", dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VAC_DEPTH), "LAYERS"),
    ("cu_on_cu_10keV", dict(tally_mask=capi.TALLY_VAC_DEPTH, potential=capi.POT_MOLIERE), "GENERIC"),
    ("uo2", dict(tally_mask=capi.TALLY_IONLOG, ionlog_z=54), "CLUSTERS-LOG"),                     # apps/mytrim_uo2
    ("uo2", dict(tally_mask=capi.TALLY_IONLOG | capi.TALLY_PHONON | capi.TALLY_RECORDS, ionlog_z=54), "CLUSTERS"),
    ("uo2", dict(tally_mask=capi.TALLY_VAC_DEPTH), "GENERIC"),
]


def setup_variant_case(e, sample):
    if sample == "uo2":
        cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
        e.set_materials([util.UO2, util.XE_GAS])
        e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
    else:
        util.setup_engine(e, sample)


@pytest.mark.parametrize("sample,cfg,want", VARIANT_CASES)
def test_configurations_select_their_kernel_variant(sample, cfg, want):
    """pick_variant() (mtb_transport.cuh; the same function the CUDA engine calls) on the BASELINE.json configurations and
    on the shapes of the in-tree drivers: each selects the leanest compile-time variant that covers it."""
    with util.HostSimEngine(**cfg) as hs:
        setup_variant_case(hs, sample)
        hs._lib.hs_variant.argtypes = [C.c_void_p]
        hs._lib.hs_variant.restype = C.c_char_p
        assert hs._lib.hs_variant(hs._h).decode() == want
