"""Parity tests proper: the CUDA path through the C ABI against the oracle with the same Philox
streams (deterministic criterion, 1e-5 relative) and against golden records of the unmodified
reference (statistical criterion).  All need a B200."""
import json
import os

import numpy as np
import pytest

from mytrim_b200 import capi
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-5  # north_star: per-ion trajectories within 1e-5 relative of the CPU replay


def _pair(cfg, name):
    eng = capi.Engine(**cfg)
    orc = util.OracleEngine(util.ORC_RNG_PHILOX, **cfg)
    c = util.setup_engine(eng, name)
    util.setup_engine(orc, name)
    return eng, orc, c


@pytest.mark.parametrize("name,n", [("cu_on_cu_10keV", 1000), ("cu_on_cu_1keV", 4000), ("h_on_fe_100keV", 1000),
                                    ("he_on_fe_100keV", 300), ("c_on_w_1MeV", 48), ("xe_on_zro2_500keV", 24)])
def test_trajectories_match_oracle(name, n):
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS | capi.TALLY_PHONON)
    eng, orc, c = _pair(cfg, name)
    ions = util.primaries_for(c, n)
    rg = eng.run(ions, seed=2344, records=True)
    ro = orc.run(ions, seed=2344, records=True)
    same = (ro["vacancies"] == rg["vacancies"]) & (ro["steps"] == rg["steps"]) & (ro["ions"] == rg["ions"])
    # fp32-vs-fp64 branch flips (Newton iteration count at |q/r| == 0.001, threshold tests) are rare but real.
    # Measured on the B200 (round 2): >= 99.9 % of the cascades identical in every configuration
    assert n - same.sum() <= max_flipped_cascades(n, int(ro["steps"].sum())), same.mean()
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    sel = ro["primary_steps"] == rg["primary_steps"]
    assert sel.mean() >= 0.995 - 1.0 / n
    rel = (np.linalg.norm(ro["pos"] - rg["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= max(1, 0.002 * len(rel)), ((rel < TOL).mean(), rel.max())
    assert np.median(rel) < 0.1 * TOL, np.median(rel)
    bad = np.abs(ro["Eel"][same] - rg["Eel"][same]) > TOL * ro["Eel"][same]
    assert bad.sum() <= max(2, 0.005 * same.sum()), bad.sum()
    cg, co = eng.counters(), orc.counters()
    E0 = c["ion"][2] * n
    assert abs(cg["EelTotal"] + cg["EnucTotal"] - E0) < 1e-6 * E0   # energy partition closes
    for k in ("vacancies_created", "replacements", "steps", "ions"):
        assert abs(cg[k] - co[k]) <= 0.02 * co[k], k
    vg, rpg = eng.vac_depth()
    vo, rpo = orc.vac_depth()
    assert cg["hist_clamped"] == 0 and vg.sum() <= cg["vacancies_created"]
    m = max(len(vg), len(vo))
    d = np.abs(np.pad(vg, (0, m - len(vg))).astype(int) - np.pad(vo, (0, m - len(vo))).astype(int)).sum()
    # cascades that took a different branch somewhere are statistically equivalent but place their
    # vacancies elsewhere: allow their share of the histogram to differ completely
    assert d <= (2.0 * (1.0 - same.mean()) + 0.01) * vo.sum() + 5
    eng.close()


# Measured levels (B200, round 2; the numbers printed by tools/per_ion_parity.py, profiles/r02_per_ion_parity.log) are
# written next to the thresholds below.
PER_ION_MIN_IDENTICAL = 0.9999     # ions whose integer fields (primary, Z, generation, tag, final state) are bit-identical
PER_ION_MAX_POS_OUTLIERS = 5e-4    # share of ions whose birth/death point is > TOL (of the distance from the source) off
# Measured (profiles/r02_per_ion_parity.log, 1 710 594 ions over the sixteen cases): against the FP32 replay 1 710 592 ions
# joined, 1 710 591 with identical integer fields (the three others: a CUT-boundary flip in the wire and one ion of the
# layer stack), 154 position outliers (9.0e-5; worst case 26 of 62 579 = 4.2e-4 on the 10 MeV Xe tracks), 22 energy
# outliers; against the FP64 oracle 1 710 591 identical, 107 position outliers (6.3e-5), 12 energy outliers.
FLIPS_PER_STEP = 3e-6              # threshold tests that flip between two arithmetics, per collision step (measured ~1e-6)


def max_flipped_cascades(n, steps_total):
    """Cascades that may differ in an integer field: FLIPS_PER_STEP x steps per cascade of them (+1 for small n)."""
    return max(0.01, FLIPS_PER_STEP * steps_total / n) * n + 1



@pytest.mark.parametrize("name,n", __import__("tests.parity_cases", fromlist=["x"]).PER_ION_CASES)
def test_per_ion_parity_with_fp32_host_replay(name, n):
    """The deterministic criterion as north_star states it: per-ION trajectories of the CUDA kernels against a CPU
    replay that uses the same Philox streams — tests/libhostsim.so, the device lane loop compiled for the host in
    FP32 (same source, same ion ids).  Every followed ion is compared through MTB_TALLY_IONLOG (birth and death
    position and energy, primary, Z, generation, tag, final state), joined by the scheduling-independent ion id; the
    FP64 oracle is held to the same level.  The two FP32 paths still differ in the last bits (MUFU.EX2/LG2/RCP/RSQ
    on the device, libm on the host; FMA contraction), so a threshold test can flip; the allowances are the
    measured levels, not slack."""
    from tests import parity_cases
    cfg = dict(tally_mask=capi.TALLY_IONLOG | capi.TALLY_RECORDS, ionlog_capacity=1 << 21, **parity_cases.case_options(name))
    with capi.Engine(**cfg) as eng, util.HostSimEngine(**cfg) as hs, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        ions = parity_cases.setup_case(eng, name, n)
        parity_cases.setup_case(hs, name, n)
        parity_cases.setup_case(orc, name, n)
        rg = eng.run(ions, seed=2344, records=True)
        rh = hs.run(ions, seed=2344, records=True)
        ro = orc.run(ions, seed=2344, records=True)
        lg = eng.ion_log(1 << 21)
        sh = util.compare_ion_logs(lg, hs.ion_log(1 << 21), ions)
        so = util.compare_ion_logs(lg, orc.ion_log(1 << 21), ions)
    for partner, s, r in (("fp32 replay", sh, util.compare_records(rg, rh, ions)),
                          ("fp64 oracle", so, util.compare_records(rg, ro, ions))):
        print("%s vs %s: %s %s" % (name, partner, s, r))
        assert s["joined"] >= PER_ION_MIN_IDENTICAL * max(s["n_test"], s["n_replay"]), (partner, s)
        assert s["ints_equal"] >= PER_ION_MIN_IDENTICAL * max(s["n_test"], s["n_replay"]), (partner, s)
        assert s["pos_outliers"] <= PER_ION_MAX_POS_OUTLIERS * s["joined"] + 2, (partner, s)
        assert s["energy_outliers"] <= 2e-4 * s["joined"] + 3, (partner, s)   # measured worst: 8 of 72 614 (C-Kr, 1 keV)
        assert s["median_rel_pos"] < 0.1 * TOL, (partner, s)
        assert r["n"] - r["cascades_identical"] <= max_flipped_cascades(r["n"], int(rg["steps"].sum())), (partner, r)
        assert r["pos_outliers"] <= 0.002 * r["n"] + 1, (partner, r)


@pytest.mark.parametrize("name,n", [("cu_on_cu_10keV", 4000), ("h_on_fe_100keV", 4000), ("c_on_w_1MeV", 64),
                                    ("xe_on_zro2_500keV", 24), ("cu_on_cu_150keV", 64), ("xe_on_uo2_10MeV", 2)])
def test_north_star_kernels_against_fp32_host_replay(name, n):
    """The ion log is produced by the option-carrying variants; the kernels bench.py times (MONO for the single-element
    samples, FAST for the compounds: TrimVacCount tallies + records only) are compared per cascade with the same variant
    of the FP32 host replay: integer record fields, end point of the primary, electronic loss, and the depth histograms."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as eng, util.HostSimEngine(**cfg) as hs:
        c = util.setup_engine(eng, name)
        util.setup_engine(hs, name)
        ions = util.primaries_for(c, n)
        rg = eng.run(ions, seed=2344, records=True)
        rh = hs.run(ions, seed=2344, records=True)
        r = util.compare_records(rg, rh, ions)
        cg, ch = eng.counters(), hs.counters()
        vg, rpg = eng.vac_depth()
        vh, rph = hs.vac_depth()
    print("%s MONO/FAST vs fp32 replay: %s" % (name, r))
    assert n - r["cascades_identical"] <= max_flipped_cascades(n, cg["steps"]), r
    assert r["pos_outliers"] <= 0.002 * n + 1 and r["eel_outliers"] <= 0.002 * n + 1 and r["median_rel_pos"] < 0.1 * TOL, r
    for k in ("vacancies_created", "replacements", "steps", "ions"):
        assert abs(cg[k] - ch[k]) <= 2e-3 * ch[k], (k, cg[k], ch[k])
    m = max(len(vg), len(vh))
    d = np.abs(np.pad(vg, (0, m - len(vg))).astype(int) - np.pad(vh, (0, m - len(vh))).astype(int)).sum()
    # the depth bin is the truncated x coordinate: an end point within the position error (~1e-3 A at 1e4 A depth) of
    # an integer lands in the neighbouring bin (measured 82 of 49 527 on the 10 MeV Xe tracks)
    assert d <= (2.0 * (1.0 - r["cascades_identical"] / n) + 5e-3) * vh.sum() + 5, d


def test_stack_of_different_materials():
    """Layer look-up over DIFFERENT materials (Cu / Fe / W / ZrO2: the FAST kernel's binary search over the cumulative
    thicknesses, compound target pick) against the oracle, whose look-up is pinned against the reference on the same
    stack (tests/test_oracle_golden.py)."""
    from tests.golden.make_golden import STACK_CASE as o
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as eng, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        for e in (eng, orc):
            util.setup_engine(e, o)
        ions = util.primaries_for(o, 600)
        rg = eng.run(ions, seed=5, records=True)
        ro = orc.run(ions, seed=5, records=True)
        cg, co = eng.counters(), orc.counters()
    same = (ro["vacancies"] == rg["vacancies"]) & (ro["steps"] == rg["steps"]) & (ro["ions"] == rg["ions"])
    assert len(ions) - same.sum() <= max_flipped_cascades(len(ions), int(ro["steps"].sum())), same.mean()
    sel = ro["primary_steps"] == rg["primary_steps"]
    assert sel.mean() >= 0.99
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    rel = (np.linalg.norm(ro["pos"] - rg["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= max(1, 0.002 * len(rel)), ((rel < TOL).mean(), rel.max())
    assert np.median(rel) < 0.1 * TOL, np.median(rel)
    for k in ("vacancies_created", "replacements", "steps", "ions"):
        assert abs(cg[k] - co[k]) <= 0.02 * co[k], k


def test_stopping_matches_oracle():
    data = json.load(open(os.path.join(util.GOLDEN, "stopping.json")))
    for name, d in data.items():
        E = np.logspace(1.5, 8.5, 400)
        with capi.Engine() as eng, util.OracleEngine(util.ORC_RNG_PHILOX) as orc:
            eng.set_materials([d["material"]])
            orc.set_materials([d["material"]])
            a = orc.stopping(0, d["Z"], d["m"], E)
            b = eng.stopping(0, np.full(len(E), d["Z"]), np.full(len(E), d["m"]), E)
            # golden known answers of the compiled reference
            g = eng.stopping(0, np.full(len(d["E"]), d["Z"]), np.full(len(d["E"]), d["m"]), d["E"])
        assert np.abs(b / a - 1).max() < TOL, (name, np.abs(b / a - 1).max())
        assert np.abs(g / np.array(d["getrstop"]) - 1).max() < TOL, name


def test_sharding_invariance():
    """Results do not depend on how primaries are split across launches / GPUs (global Philox ids)."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    c = util.CONFIGS["cu_on_cu_10keV"]
    ions = util.primaries_for(c, 3000)
    with capi.Engine(**cfg) as a, capi.Engine(**cfg) as b:
        util.setup_engine(a, c)
        util.setup_engine(b, c)
        ra = a.run(ions, seed=77, records=True)
        rb = np.concatenate([b.run(ions[:1100], seed=77, first_index=0, records=True),
                             b.run(ions[1100:], seed=77, first_index=1100, records=True)])

        def same_records(x, y):
            for f in x.dtype.names:
                if f in ("Eel", "Enuc"):   # sums of per-lane partial sums: order of addition is not fixed
                    assert np.allclose(x[f], y[f], rtol=1e-12, atol=0), f
                else:
                    assert np.array_equal(x[f], y[f]), f

        same_records(ra, rb)
        va, _ = a.vac_depth()
        vb, _ = b.vac_depth()
        assert np.array_equal(va, vb)
        ca, cb = a.counters(), b.counters()
        assert ca["steps"] == cb["steps"] and ca["vacancies_created"] == cb["vacancies_created"]
        # beam mode (template ion) is the same thing without the per-primary host array
        b.reset_tallies()
        rc = b.run_beam(3000, ions[0], seed=77, records=True)
        same_records(ra, rc)


def test_follow_policies_and_vacancy_models():
    for cfg in (dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT, tally_mask=capi.TALLY_RANGE),
                dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP),
                dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=1, vacancy_model=capi.VAC_KP),
                dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VACMAP, vmap_z=(29, 8, -1))):
        eng, orc, c = _pair(cfg, "cu_on_cu_10keV")
        ions = util.primaries_for(c, 600)
        eng.run(ions, seed=11)
        orc.run(ions, seed=11)
        cg, co = eng.counters(), orc.counters()
        for k in ("steps", "ions", "replacements", "recoils_queued", "vacancies_created"):
            assert abs(cg[k] - co[k]) <= 0.01 * co[k] + 1, (cfg, k, cg[k], co[k])
        if cfg.get("tally_mask", 0) & capi.TALLY_RANGE:
            xg, zg = eng.range_list()
            xo, zo = orc.range_list()
            assert abs(len(xg) - len(xo)) <= 0.01 * len(xo)
            assert abs(np.mean(xg) - np.mean(xo)) < 0.02 * abs(np.mean(xo))
        if cfg.get("tally_mask", 0) & capi.TALLY_VAC_ENERGY:
            eg, eo = eng.vac_energy(), orc.vac_energy()
            assert abs(int(eg.sum()) - int(eo.sum())) <= 0.01 * eo.sum()
            assert np.abs(eg.sum(axis=1).astype(int) - eo.sum(axis=1).astype(int)).sum() <= 0.02 * eo.sum()
            assert abs(int(eng.vacmap().sum()) - int(orc.vacmap().sum())) <= 0.01 * orc.vacmap().sum()
        eng.close()


@pytest.mark.parametrize("name,n,cfg", [
    ("cu_on_cu_10keV", 1500, dict(tally_mask=capi.TALLY_VAC_DEPTH)),                                   # MONO
    ("c_on_w_1MeV", 150, dict(tally_mask=capi.TALLY_VAC_ENERGY)),                                      # MONO-EVAC
    ("xe_on_zro2_500keV", 40, dict(tally_mask=capi.TALLY_PHONON)),                                     # FAST-PHONON
    ("cu_on_cu_10keV", 1500, dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VAC_DEPTH | capi.TALLY_VACMAP,
                                  vmap_z=(29, 8, -1))),                                                # LAYERS
    ("cu_on_cu_10keV", 3000, dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT, tally_mask=capi.TALLY_RANGE)),
    ("xe_on_zro2_500keV", 400, dict(follow=capi.FOLLOW_GEN_LT, follow_max_gen=2, vacancy_model=capi.VAC_KP)),
])
def test_tally_bins_against_fp32_host_replay(name, n, cfg):
    """Every tally of the in-tree classes BIN BY BIN (not by sums) against the FP32 host replay of the device loop on the
    same primaries and Philox streams: depth histograms of TrimVacCount, the 2-D histogram of TrimVacEnergyCount, the
    TrimVacMap grid, TrimRange's list, TrimPhononOut's energy partition and the Kinchin-Pease counters of
    TrimPrimaries/TrimRecoils.  The two sides differ by the rare threshold flips of the two FP32 arithmetics
    (~1e-6 per collision step; a flip moves the events of one sub-cascade, hundreds of them in a
    1 MeV cascade in tungsten): the summed absolute difference over all bins may reach 2e-3 of the tallied events (measured:
    1.1e-3 on C->W, less elsewhere) — never a systematic shift of a bin, which would show up as O(1)."""
    with capi.Engine(**cfg) as eng, util.HostSimEngine(**cfg) as hs:
        for e in (eng, hs):
            c = util.setup_engine(e, name)
        ions = util.primaries_for(c, n)
        eng.run(ions, seed=41)
        hs.run(ions, seed=41)
        cg, ch = eng.counters(), hs.counters()
        steps = ch["steps"]
        allow = 16 + 1e-4 * steps           # events that may appear / vanish with a flipped sub-cascade
        for k in ("steps", "ions", "replacements", "recoils_queued", "vacancies_created"):
            assert abs(cg[k] - ch[k]) <= allow * (10 if k == "steps" else 1), (k, cg[k], ch[k])
        mask = cfg.get("tally_mask", 0)
        if mask & capi.TALLY_VAC_DEPTH:
            for hg, hh in zip(eng.vac_depth(), hs.vac_depth()):
                m = max(len(hg), len(hh))
                hg, hh = np.pad(hg, (0, m - len(hg))).astype(np.int64), np.pad(hh, (0, m - len(hh))).astype(np.int64)
                assert hh.sum() > 0 and np.abs(hg - hh).sum() <= 16 + 2e-3 * hh.sum(), (np.abs(hg - hh).sum(), hh.sum())
        if mask & capi.TALLY_VAC_ENERGY:
            eg, eh = eng.vac_energy(rows=32, bins=16384).astype(np.int64), hs.vac_energy(rows=32, bins=16384).astype(np.int64)
            assert eh.sum() > 0 and np.abs(eg - eh).sum() <= 16 + 2e-3 * eh.sum(), (np.abs(eg - eh).sum(), eh.sum())
        if mask & capi.TALLY_VACMAP:
            vg, vh = eng.vacmap().astype(np.int64), hs.vacmap().astype(np.int64)
            assert vh.sum() > 0 and np.abs(vg - vh).sum() <= 16 + 2e-3 * vh.sum(), (np.abs(vg - vh).sum(), vh.sum())
        if mask & capi.TALLY_RANGE:
            (xg, zg), (xh, zh) = eng.range_list(), hs.range_list()
            assert abs(len(xg) - len(xh)) <= allow and len(xh) > 0
            if len(xg) == len(xh):
                assert np.abs(np.sort(xg) - np.sort(xh)).max() <= 1e-3 * np.abs(xh).max() or \
                    (np.abs(np.sort(xg) - np.sort(xh)) > 1e-3 * np.abs(xh).max()).sum() <= allow
        if mask & capi.TALLY_PHONON:
            E0 = ions["E"].sum()
            assert abs(cg["EelTotal"] + cg["EnucTotal"] - E0) < 1e-6 * E0
            assert abs(cg["EnucTotal"] - ch["EnucTotal"]) <= 2e-5 * E0 and abs(cg["EelTotal"] - ch["EelTotal"]) <= 2e-5 * E0


def test_ion_log_and_single_ion_events():
    cfg = dict(tally_mask=capi.TALLY_IONLOG, ionlog_z=8)
    eng, orc, c = _pair(cfg, "xe_on_zro2_500keV")
    ions = util.primaries_for(c, 2)
    eng.run(ions, seed=5)
    orc.run(ions, seed=5)
    lg, lo = eng.ion_log(), orc.ion_log()
    assert abs(len(lg) - len(lo)) <= 0.05 * len(lo)
    common = np.intersect1d(lg["uid"], lo["uid"])
    assert len(common) >= 0.9 * len(lo)
    ion = ions[0]
    fg, sg, eg = eng.trim_one(ion, 99, 1234)
    fo, so, eo = orc.trim_one(ion, 99, 1234)
    assert sg == so and len(eg) == len(eo) > 10
    for f in ("material", "element", "pka_state", "recoil_above_threshold"):
        assert np.array_equal(eg[f], eo[f]), f
    assert np.abs(eg["pka_pos"] - eo["pka_pos"]).max() < TOL * np.abs(eo["pka_pos"]).max()
    assert np.abs(fg["pos"] - fo["pos"]).max() < TOL * np.abs(fo["pos"]).max()
    eng.close()


def test_wire_and_clusters_geometry():
    """CUT boundaries / vacuum (SampleWire) and the spatial-hash cluster lookup (sampleClusters)."""
    # wire: Cu wire 200 A across, ions start on the axis and fly along z
    cfg = dict(tally_mask=capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as eng, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        for e in (eng, orc):
            e.set_materials([util.CU])
            e.set_geometry(capi.GEOM_WIRE, (200.0, 200.0, 1000.0), bc=(capi.BC_CUT, capi.BC_CUT, capi.BC_PBC))
        ions = capi.make_ions(400, 29, 63.546, 5e4, pos=(100.0, 60.0, 0.0), direction=(0.0, 0.3, 1.0))
        rg = eng.run(ions, seed=3, records=True)
        ro = orc.run(ions, seed=3, records=True)
        cg, co = eng.counters(), orc.counters()
        assert co["left_sample"] > 0
        for k in ("steps", "ions", "left_sample", "lost", "vacancies_created"):
            assert abs(cg[k] - co[k]) <= 0.02 * co[k] + 2, (k, cg[k], co[k])
        assert (rg["state"] == ro["state"]).mean() > 0.97
    # clusters: UO2 matrix with Xe bubbles (tests/uo2 geometry, 4 bubbles)
    cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_IONLOG, ionlog_z=54)
    with capi.Engine(**cfg) as eng, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        for e in (eng, orc):
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        # Xe ions starting inside bubble 0, flying out through the matrix
        ions = capi.make_ions(200, 54, 132.0, 3e4, pos=tuple(cl[0, :3]), direction=(0.6, 0.0, 0.8))
        rg = eng.run(ions, seed=9, records=True)
        ro = orc.run(ions, seed=9, records=True)
        cg, co = eng.counters(), orc.counters()
        for k in ("steps", "ions", "vacancies_created"):
            assert abs(cg[k] - co[k]) <= 0.02 * co[k] + 2, (k, cg[k], co[k])
        lg, lo = eng.ion_log(), orc.ion_log()
        # Xe recoils knocked out of a bubble carry that bubble's index as tag
        assert set(np.unique(lg["tag"])) <= {-1, 0, 1, 2, 3} and (lg["tag"] >= 0).sum() > 0
        assert abs((lg["tag"] >= 0).sum() - (lo["tag"] >= 0).sum()) <= 0.05 * (lo["tag"] >= 0).sum() + 3


@pytest.mark.parametrize("name", ["cu_on_cu_10keV", "h_on_fe_100keV", "he_on_fe_100keV"])
def test_statistics_against_reference_golden(name):
    """Two-sample KS against per-primary records of the unmodified reference (tests/golden)."""
    from scipy import stats
    gold = np.load(os.path.join(util.GOLDEN, "ref_records_%s.npz" % name))["records"]
    c = util.CONFIGS[name]
    n = 20000
    with capi.Engine(tally_mask=capi.TALLY_RECORDS) as eng:
        util.setup_engine(eng, c)
        rec = eng.run(util.primaries_for(c, n), seed=4242, records=True)
    for field, getter in (("x", lambda r: r["pos"][:, 0]), ("vacancies", lambda r: r["vacancies"].astype(float)),
                          ("Eel", lambda r: r["Eel"]),
                          ("lateral", lambda r: np.hypot(r["pos"][:, 1] - 50.0, r["pos"][:, 2] - 50.0))):
        p = stats.ks_2samp(getter(rec), getter(gold)).pvalue
        assert p > 0.001, (name, field, p)


@pytest.mark.parametrize("name", ["cu_on_cu_10keV", "cu_on_cu_1keV", "h_on_fe_100keV", "he_on_fe_100keV", "c_on_w_1MeV",
                                  "xe_on_zro2_500keV", "cu_on_cu_150keV", "h_on_fe_1MeV", "xe_on_uo2_10MeV"])
def test_north_star_statistical_criterion(name):
    """BASELINE.json's statistical criterion at full size — 1e6 Cu->Cu 10 keV cascades, and 1e5..1e6 cascades of the
    other configurations — on the GPU against as many cascades of the UNMODIFIED reference (distinct 32-bit seeds),
    summarised in tests/golden/ref_stats_<name>.npz (quantiles / exact histograms; tests/util.py::ks_against_summary):
    two-sample KS p > 0.01 and means within 1 % on projected range, lateral range, electronic loss (= energy
    partition), vacancies, replacements, collision steps and followed ions."""
    summary = np.load(os.path.join(util.GOLDEN, "ref_stats_%s.npz" % name))
    c = util.CONFIGS[name]
    n = int(summary["n"])
    # TrimVacCount tallies + records: the north-star kernels (MONO for the single-element samples, FAST for the ZrO2 stack)
    with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS) as eng:
        util.setup_engine(eng, c)
        rec = eng.run(util.primaries_for(c, n), seed=2344, records=True)
    res = util.ks_against_summary(rec, summary)
    assert len(res) == 7
    for obs, (mean, ref_mean, D, p) in res.items():
        se = np.sqrt(2.0 * float(summary["m_" + obs][1]) / n)
        # the north-star level for every configuration (the run is deterministic: fixed seed, scheduling-independent
        # streams — it either passes always or never)
        assert p > 0.01, (name, obs, D, p)
        assert abs(mean - ref_mean) <= max(0.01 * abs(ref_mean), 4.0 * se), (name, obs, mean, ref_mean)


def test_multi_gpu_allreduce_in_process():
    """mtb_allreduce: two handles on two GPUs, primaries sharded by global index, NCCL tally reduction
    equals one GPU running everything (needs >= 2 GPUs)."""
    import ctypes as C
    lib = capi.load_library()
    if lib.mtb_device_count() < 2:
        pytest.skip("needs two GPUs")
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH)
    c = util.CONFIGS["cu_on_cu_10keV"]
    n = 20000
    ions = util.primaries_for(c, n)
    with capi.Engine(device=0, **cfg) as one, capi.Engine(device=0, **cfg) as a, capi.Engine(device=1, **cfg) as b:
        for e in (one, a, b):
            util.setup_engine(e, c)
        one.run(ions, seed=5)
        a.run(ions[:n // 2], seed=5, first_index=0)
        b.run(ions[n // 2:], seed=5, first_index=n // 2)
        arr = (C.c_void_p * 2)(a._h, b._h)
        lib.mtb_allreduce.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        assert lib.mtb_allreduce(arr, 2) == 0, lib.mtb_last_error()
        c1, ca, cb = one.counters(), a.counters(), b.counters()
        for k in ("vacancies_created", "replacements", "steps", "ions", "primaries"):
            assert c1[k] == ca[k] == cb[k], k
        assert abs(c1["EelTotal"] - ca["EelTotal"]) <= 1e-9 * c1["EelTotal"]
        assert np.array_equal(one.vac_depth()[0], a.vac_depth()[0])
        assert np.array_equal(one.vac_depth()[1], b.vac_depth()[1])


def test_per_primary_species_and_clusters():
    """Heterogeneous primaries (own (Z, m) each, like fission fragments) in the tests/uo2 geometry."""
    from tests.test_device_loop_host import _fission_like_primaries
    cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_PHONON)
    ions = _fission_like_primaries(400)
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, capi.Engine(**cfg) as eng:
        for e in (orc, eng):
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
        ro = orc.run(ions, seed=17, records=True)
        rg = eng.run(ions, seed=17, records=True)
        co, cg = orc.counters(), eng.counters()
    same = (ro["vacancies"] == rg["vacancies"]) & (ro["steps"] == rg["steps"]) & (ro["ions"] == rg["ions"])
    assert same.mean() >= 0.8, same.mean()
    sel = ro["primary_steps"] == rg["primary_steps"]
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0)
    rel = (np.linalg.norm(ro["pos"] - rg["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= max(2, 0.005 * len(rel)) and np.median(rel) < 0.1 * TOL
    E0 = ions["E"].sum()
    assert abs(cg["EelTotal"] + cg["EnucTotal"] - E0) < 1e-6 * E0
    assert abs(cg["steps"] - co["steps"]) <= 0.02 * co["steps"]


def test_pinned_host_primaries_are_read_in_place():
    """mtb_run with page-locked primaries (zero-copy) gives bit-identical records to the staged path."""
    import torch
    c = util.CONFIGS["cu_on_cu_10keV"]
    n = 5000
    ions = util.primaries_for(c, n)
    pinned = torch.empty(n * capi.ION_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    view = np.frombuffer(pinned.numpy(), dtype=capi.ION_DTYPE)
    view[:] = ions
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as a, capi.Engine(**cfg) as b:
        util.setup_engine(a, c)
        util.setup_engine(b, c)
        ra = a.run(ions, seed=8, records=True)
        rb = np.zeros(n, dtype=capi.RECORD_DTYPE)
        lib = capi.load_library()
        assert lib.mtb_run(b._h, n, pinned.data_ptr(), 8, 0, rb.ctypes.data) == 0, lib.mtb_last_error()
        for f in ra.dtype.names:
            if f in ("Eel", "Enuc"):
                assert np.allclose(ra[f], rb[f], rtol=1e-12, atol=0), f
            else:
                assert np.array_equal(ra[f], rb[f]), f


def test_fast_kernel_defers_unknown_species():
    """The compile-time fast kernel hands primaries without a projectile class to the generic kernel;
    the union is bit-identical (per-primary records) to running everything through the generic kernel."""
    from tests.test_device_loop_host import _fission_like_primaries
    ions = _fission_like_primaries(3000)
    ions["pos"] = (0.0, 50.0, 50.0)
    ions["dir"] = (1.0, 0.0, 0.0)
    ions[::3] = capi.make_ions(1000, 29, 63.546, 1.0e4)
    fast = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)              # -> TraitsFast + deferral
    generic = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS | capi.TALLY_PHONON)  # -> TraitsGeneric
    with capi.Engine(**fast) as a, capi.Engine(**generic) as b:
        util.setup_engine(a, "cu_on_cu_10keV")
        util.setup_engine(b, "cu_on_cu_10keV")
        ra = a.run(ions, seed=4, records=True)
        rb = b.run(ions, seed=4, records=True)
        # Primaries with a projectile class (the Cu ions and the first 16 distinct species, which the
        # host registers) run in the fast instantiation in `a` and in the generic one in `b`; the rest
        # are deferred to the generic kernel in `a`.  Different instantiations contract FMAs
        # differently, so agreement is to the trajectory tolerance, not bitwise.
        same = (ra["steps"] == rb["steps"]) & (ra["vacancies"] == rb["vacancies"]) & (ra["ions"] == rb["ions"])
        assert same.mean() > 0.97, same.mean()
        sel = ra["primary_steps"] == rb["primary_steps"]
        d = np.linalg.norm(ra["pos"] - rb["pos"], axis=1) / np.maximum(np.linalg.norm(ra["pos"] - ions["pos"], axis=1), 1.0)
        assert (d[sel] >= TOL).sum() <= 0.005 * len(ions) + 2
        assert (ra["state"] == rb["state"]).mean() > 0.99
        ca, cb = a.counters(), b.counters()
        assert abs(ca["steps"] - cb["steps"]) <= 1e-3 * cb["steps"] and ca["primaries"] == cb["primaries"] == len(ions)


@pytest.mark.parametrize("which", ["mono", "fast", "clusters", "layers", "clusters_log", "mono_evac", "fast_phonon"])
def test_lean_variants_agree_with_generic_kernel(which):
    """Every lean kernel variant (mtb_transport.cuh: MONO / FAST / CLUSTERS / LAYERS / CLUSTERS-LOG / MONO-EVAC) against the all-options
    kernel on the same primaries and seeds.  Different instantiations contract FMAs differently, so agreement is
    to the trajectory tolerance; the integer tallies agree for all cascades without a branch flip.  (A
    single-element sample takes MONO by default; MYTRIM_B200_NO_MONO routes it through FAST.)"""
    knob = {"fast": "MYTRIM_B200_NO_MONO"}.get(which)
    if which in ("mono", "fast"):
        cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
        def setup(e):
            c = util.setup_engine(e, "cu_on_cu_10keV")
            return util.primaries_for(c, 4000)
    elif which == "mono_evac":
        # validation/c_on_w/input.json: TrimVacEnergyCount on a single-element sample
        cfg = dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_RECORDS)
        def setup(e):
            c = util.setup_engine(e, "c_on_w_1MeV")
            return util.primaries_for(c, 1500)
    elif which == "fast_phonon":
        # TrimPhononOut's energy partition on a layered compound sample
        cfg = dict(tally_mask=capi.TALLY_PHONON | capi.TALLY_RECORDS)
        def setup(e):
            c = util.setup_engine(e, "xe_on_zro2_500keV")
            return util.primaries_for(c, 600)
    elif which == "layers":
        cfg = dict(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS, follow=capi.FOLLOW_GEN_LT,
                   follow_max_gen=2, vacancy_model=capi.VAC_KP)
        def setup(e):
            c = util.setup_engine(e, "xe_on_zro2_500keV")
            return util.primaries_for(c, 600)
    else:
        # "clusters": ion log + energy partition (run-time tallies); "clusters_log": the tests/uo2 driver's own mask
        cfg = dict(tally_mask=(capi.TALLY_PHONON if which == "clusters" else 0) | capi.TALLY_RECORDS | capi.TALLY_IONLOG, ionlog_z=54)
        cl = np.loadtxt(os.path.join(util.GOLDEN, "uo2_out.clcoor"))[:, :4]
        def setup(e):
            from tests.test_device_loop_host import _fission_like_primaries
            e.set_materials([util.UO2, util.XE_GAS])
            e.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
            ions = _fission_like_primaries(400)
            ions["pos"][:40] = cl[np.arange(40) % len(cl), :3] + 2.0
            return ions
    if knob:
        os.environ[knob] = "1"
    try:
        with capi.Engine(**cfg) as a:
            ions = setup(a)
            ra = a.run(ions, seed=31, records=True)
            ca = a.counters()
            la = len(a.ion_log()) if which.startswith("clusters") else 0
            ea = a.vac_energy(rows=32, bins=16384) if which == "mono_evac" else None
    finally:
        if knob:
            os.environ.pop(knob)
    os.environ["MYTRIM_B200_VARIANT"] = "generic"
    try:
        with capi.Engine(**cfg) as b:
            setup(b)
            rb = b.run(ions, seed=31, records=True)
            cb = b.counters()
            lb = len(b.ion_log()) if which.startswith("clusters") else 0
            eb = b.vac_energy(rows=32, bins=16384) if which == "mono_evac" else None
    finally:
        os.environ.pop("MYTRIM_B200_VARIANT")
    same = (ra["steps"] == rb["steps"]) & (ra["vacancies"] == rb["vacancies"]) & (ra["ions"] == rb["ions"])
    assert same.mean() > (0.97 if which in ("mono", "fast") else 0.85), same.mean()
    sel = ra["primary_steps"] == rb["primary_steps"]
    assert sel.mean() > 0.97
    d = np.linalg.norm(ra["pos"] - rb["pos"], axis=1) / np.maximum(np.linalg.norm(ra["pos"] - ions["pos"], axis=1), 1.0)
    assert (d[sel] >= TOL).sum() <= 0.005 * len(ions) + 2
    assert abs(ca["steps"] - cb["steps"]) <= 2e-3 * cb["steps"] and ca["primaries"] == cb["primaries"] == len(ions)
    assert abs(ca["vacancies_created"] - cb["vacancies_created"]) <= 2e-3 * cb["vacancies_created"]
    assert abs(la - lb) <= 0.02 * lb + 4
    if which == "fast_phonon":
        E0 = ions["E"].sum()
        assert abs(ca["EelTotal"] + ca["EnucTotal"] - E0) < 1e-6 * E0
        assert abs(ca["EnucTotal"] - cb["EnucTotal"]) < 2e-3 * cb["EnucTotal"]
    if ea is not None:
        # the 2-D tally of TrimVacEnergyCount: row sums (energy decades) and depth profile agree to the flip level
        assert ea.sum() > 0
        assert abs(int(ea.sum()) - int(eb.sum())) <= 2e-3 * eb.sum()
        assert np.abs(ea.sum(axis=1).astype(float) - eb.sum(axis=1)).max() <= 5e-3 * eb.sum(axis=1).max() + 4


def test_work_sharing_pool_is_result_neutral():
    """Few large cascades: idle lanes adopt suspended ions from the shared pool.  Per-ion Philox streams
    make the result independent of which lane follows which ion: integer tallies and per-primary
    counters are identical with and without sharing, energies agree to rounding."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    for name, n in (("c_on_w_1MeV", 64), ("xe_on_zro2_500keV", 40), ("cu_on_cu_10keV", 3000)):
        c = util.CONFIGS[name]
        ions = util.primaries_for(c, n)
        os.environ["MYTRIM_B200_NO_SHARE"] = "1"
        try:
            with capi.Engine(**cfg) as solo:
                util.setup_engine(solo, c)
                r0 = solo.run(ions, seed=12, records=True)
                c0, h0, ms0 = solo.counters(), solo.vac_depth(), solo.last_kernel_ms()
        finally:
            os.environ.pop("MYTRIM_B200_NO_SHARE")
        with capi.Engine(**cfg) as shared:
            util.setup_engine(shared, c)
            r1 = shared.run(ions, seed=12, records=True)
            c1, h1, ms1 = shared.counters(), shared.vac_depth(), shared.last_kernel_ms()
        for f in ("vacancies", "replacements", "steps", "ions", "state", "primary_steps"):
            assert np.array_equal(r0[f], r1[f]), (name, f)
        assert np.array_equal(r0["pos"], r1["pos"]) and np.array_equal(r0["E"], r1["E"])
        assert np.allclose(r0["Eel"], r1["Eel"], rtol=1e-12)
        for k in ("vacancies_created", "replacements", "steps", "ions", "primaries", "recoils_queued"):
            assert c0[k] == c1[k], (name, k)
        assert np.array_equal(h0[0], h1[0]) and np.array_equal(h0[1], h1[1])
        print("%s n=%d: %.2f ms without sharing, %.2f ms with (x%.1f)" % (name, n, ms0, ms1, ms0 / ms1))


@pytest.mark.parametrize("opts", [dict(potential=capi.POT_MOLIERE), dict(potential=capi.POT_CKR),
                                  dict(length_scale=10.0), dict(tmin=1.0, cw=0.01)])
def test_options_potentials_and_scale(opts):
    """MOLIERE / C-Kr potentials (trim.C:206-222, 247-259), SimconfType::setLengthScale, tmin/cw, and
    per-element Edisp/Elbind + a non-default final energy, against the oracle."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS, **opts)
    mat = {"rho": 8.92, "elements": [{"Z": 29, "m": 63.546, "t": 1.0, "Edisp": 30.0, "Elbind": 2.0}]}
    scale = opts.get("length_scale", 1.0)
    with util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc, capi.Engine(**cfg) as eng:
        for e in (orc, eng):
            e.set_materials([mat])
            e.set_layers([1000.0 / scale], wy=100.0 / scale, wz=100.0 / scale)
        ions = capi.make_ions(1500, 29, 63.546, 2.0e4, pos=(0.0, 50.0 / scale, 50.0 / scale), Ef=5.0)
        ro = orc.run(ions, seed=21, records=True)
        rg = eng.run(ions, seed=21, records=True)
        co, cg = orc.counters(), eng.counters()
    same = (ro["vacancies"] == rg["vacancies"]) & (ro["steps"] == rg["steps"]) & (ro["ions"] == rg["ions"])
    assert same.mean() >= 0.9, same.mean()
    sel = ro["primary_steps"] == rg["primary_steps"]
    path = np.maximum(np.linalg.norm(ro["pos"] - ions["pos"], axis=1), 1.0 / scale)
    rel = (np.linalg.norm(ro["pos"] - rg["pos"], axis=1) / path)[sel]
    assert (rel >= TOL).sum() <= max(2, 0.005 * len(rel)) and np.median(rel) < 0.1 * TOL
    assert abs(co["vacancies_created"] - cg["vacancies_created"]) <= 0.01 * co["vacancies_created"]


def test_published_vacancies_per_ion():
    """The reference's only published numbers for this path: validation/vacancy_count/
    vacancy_count_comparison.dat, "MyTRIM: exact" (full cascades) and "MyTRIM: KP" (TrimRange:
    primaries only + NRT estimate), 1 and 10 keV rows (fixture: tests/golden/vacancy_count_published.json)."""
    pub = json.load(open(os.path.join(util.GOLDEN, "vacancy_count_published.json")))
    targets = {"cu_on_cu": (util.CU, (29, 63.546)),
               "xe_on_u": ({"rho": 19.05, "elements": [{"Z": 92, "m": 238.03, "t": 1.0}]}, (54, 131.3)),
               "si_on_c": ({"rho": 3.51, "elements": [{"Z": 6, "m": 12.011, "t": 1.0}]}, (14, 28.086))}
    for key, (mat, (Z, m)) in targets.items():
        for row in (0, 1):
            E = pub["energy_keV"][row] * 1e3
            n = 40000
            for mode, cfg in (("exact", dict()),
                              ("kp", dict(follow=capi.FOLLOW_NONE, vacancy_model=capi.VAC_NRT))):
                with capi.Engine(**cfg) as eng:
                    eng.set_materials([mat])
                    eng.set_layers([100000.0])
                    eng.run(capi.make_ions(n, Z, m, E), seed=31)
                    vpi = eng.counters()["vacancies_created"] / n
                want = pub["%s_%s" % (key, mode)][row]
                # the published table was produced with unknown densities/masses for U and C: 4 %
                assert abs(vpi - want) < 0.04 * want, (key, mode, E, vpi, want)


def test_edge_cases_empty_single_and_degenerate_primaries():
    """Empty batch, one primary, and primaries the reference treats specially: zero energy (the reference
    would divide by zero; the engine parks the ion as an interstitial without a collision), a start in
    front of the first layer and beyond the last one (sample_layers.C:26-49: first / last layer), an
    energy below the displacement threshold."""
    cfg = dict(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as eng, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        c = util.setup_engine(eng, "cu_on_cu_10keV")
        util.setup_engine(orc, "cu_on_cu_10keV")
        rec = eng.run(util.primaries_for(c, 0), seed=3, records=True)
        assert len(rec) == 0
        cnt = eng.counters()
        assert cnt["primaries"] == 0 and cnt["steps"] == 0 and cnt["vacancies_created"] == 0
        assert eng.vac_depth()[0].sum() == 0
        one = util.primaries_for(c, 1)
        r1, o1 = eng.run(one, seed=3, records=True), orc.run(one, seed=3, records=True)
        assert r1["steps"][0] == o1["steps"][0] and r1["vacancies"][0] == o1["vacancies"][0]
        eng.reset_tallies()
        ions = util.primaries_for(c, 5)
        ions["E"][0] = 0.0
        ions["pos"][1] = (-40.0, 50.0, 50.0)
        ions["pos"][2] = (5000.0, 50.0, 50.0)
        ions["E"][3] = 10.0
        ions["dir"][4] = (0.0, 0.0, 2.0)     # not normalised, along z
        r = eng.run(ions, seed=5, records=True)
        o = orc.run(ions[1:], seed=5, first_index=1, records=True)
        assert r["steps"][0] == 0 and r["state"][0] == capi.INTERSTITIAL and r["E"][0] == 0.0
        assert np.array_equal(r["steps"][1:], o["steps"]) and np.array_equal(r["vacancies"][1:], o["vacancies"])
        assert np.array_equal(r["state"][1:], o["state"])
        d = np.abs(r["pos"][1:] - o["pos"]).max(axis=1)
        assert (d < 1e-5 * np.maximum(np.abs(o["pos"]).max(axis=1), 1.0)).all()
        assert r["vacancies"][3] == 0 and r["ions"][3] == 1


def test_error_paths_return_status_codes():
    """Nothing exits or throws across the C ABI: bad input comes back as a status + message."""
    import ctypes as C
    lib = capi.load_library()
    with capi.Engine() as eng:
        with pytest.raises(capi.MytrimError) as e:
            eng.run(capi.make_ions(10, 29, 63.546, 1e4), seed=1)          # no materials yet
        assert e.value.code == capi.EINVAL and "materials" in str(e.value)
        with pytest.raises(capi.MytrimError):
            eng.set_materials([{"rho": 8.92, "elements": [{"Z": 120, "m": 300.0, "t": 1.0}]}])  # Z > 92
        with pytest.raises(capi.MytrimError):
            eng.set_materials([{"rho": -1.0, "elements": [{"Z": 29, "m": 63.5, "t": 1.0}]}])
        with pytest.raises(capi.MytrimError):
            eng.set_geometry(capi.GEOM_LAYERS, (10.0, 10.0, 10.0))          # layers geometry without layers
        with pytest.raises(capi.MytrimError):
            eng.set_geometry(7, (10.0, 10.0, 10.0))
        eng.set_materials([util.CU])
        eng.set_layers([1000.0])
        with pytest.raises(capi.MytrimError) as e:
            eng.vac_energy()                                                 # tally not enabled
        assert e.value.code == capi.EINVAL
        with pytest.raises(capi.MytrimError) as e:
            eng.stopping(3, [29], [63.5], [1e4])                             # material index out of range
        assert e.value.code == capi.EINVAL
        # an event buffer that is too small reports the needed size
        ion = capi.make_ions(1, 29, 63.546, 1e4)
        ev = np.zeros(2, dtype=capi.EVENT_DTYPE)
        n = C.c_size_t()
        st = C.c_int32()
        rc = lib.mtb_trim_one(eng._h, ion.ctypes.data, 1, 1, C.byref(st), ev.ctypes.data, 2, C.byref(n))
        assert rc == capi.ECAPACITY and n.value > 2
        # primaries the tables cannot describe are skipped on the device and reported: the good
        # ones of the batch are followed, the handle stays usable
        bad = capi.make_ions(64, 29, 63.546, 1e4)
        bad["Z"][3] = 0
        bad["Z"][7] = 200
        bad["m"][11] = -1.0
        bad["E"][13] = np.nan
        bad["dir"][17] = 0.0
        with pytest.raises(capi.MytrimError) as e:
            eng.run(bad, seed=1)
        assert e.value.code == capi.EINVAL and "5 primaries were skipped" in str(e.value)
        assert eng.counters()["primaries"] == 59
        eng.run(capi.make_ions(64, 29, 63.546, 1e4), seed=1)
        assert eng.counters()["primaries"] == 59 + 64
        with pytest.raises(capi.MytrimError):
            eng.run_beam(8, capi.make_ions(1, 0, 63.546, 1e4)[0], seed=1)     # template ion with Z = 0
        with pytest.raises(capi.MytrimError) as e:
            eng.trim_one(capi.make_ions(1, 200, 63.546, 1e4)[0], 1, 1)        # event mode validates too
        assert e.value.code == capi.EINVAL
        with pytest.raises(capi.MytrimError) as e:
            eng.trim_many(bad[:20], 1, 0, 16)
        assert e.value.code == capi.EINVAL
        eng.trim_one(capi.make_ions(1, 29, 63.546, 1e4)[0], 1, 1)             # and the handle stays usable
    cfg = capi.default_config(potential=9)
    h = C.c_void_p()
    assert lib.mtb_create(C.byref(cfg), C.byref(h)) == capi.EINVAL
    cfg = capi.default_config(device=99)
    assert lib.mtb_create(C.byref(cfg), C.byref(h)) == capi.EINVAL


def test_engines_with_different_table_sizes_coexist():
    """The dynamic shared-memory limit is an attribute of the kernel, not of a handle: an engine with small tables
    created while one with histogram mirrors is alive must not shrink the limit under it (the façade keeps a batch
    and a single-ion engine per Trim object; bench.py one engine per configuration)."""
    c = util.CONFIGS["cu_on_cu_10keV"]
    ions = util.primaries_for(c, 2000)
    with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS) as big:
        util.setup_engine(big, c)
        r0 = big.run(ions, seed=3, records=True)
        with capi.Engine(tally_mask=0) as small:
            util.setup_engine(small, c)
            small.run(ions[:100], seed=3)
            ion, state, ev = small.trim_one(ions[0], 3, 7)
            assert len(ev) > 0
            r1 = big.run(ions, seed=3, records=True)
    assert np.array_equal(r0["steps"], r1["steps"]) and np.array_equal(r0["pos"], r1["pos"])


def test_trim_many_equals_trim_one_per_ion():
    """mtb_trim_many (one launch, one lane per ion: the façade's hand-over of a whole generation of queued ions) reports
    for every ion exactly the events mtb_trim_one reports for it with the same stream id; ions with more collisions
    than the buffer holds report their count and a valid prefix."""
    from tests import parity_cases
    cfg = dict(tally_mask=0)
    with capi.Engine(**cfg) as eng:
        ions = parity_cases.setup_case(eng, "uo2_fission_like", 70)
        ions["E"][:6] = [30.0, 80.0, 300.0, 2e3, 5e4, 3.0]      # very short and long trajectories in one batch
        K = 64
        fin, st, cnt, ev = eng.trim_many(ions, 99, 1000, K)
        assert (cnt > 0).all() and (cnt > K).any() and (cnt <= K).any()
        for i in range(len(ions)):
            f1, s1, e1 = eng.trim_one(ions[i], 99, 1000 + i, capacity=1 << 16)
            assert len(e1) == cnt[i], i
            m = min(K, int(cnt[i]))
            assert ev[i, :m].tobytes() == e1[:m].tobytes(), i      # bit-identical event records
            if cnt[i] <= K:
                assert st[i] == s1 and fin[i].tobytes() == f1.tobytes(), i
        # the long ones again with explicit stream ids and a buffer of the reported size
        longer = np.nonzero(cnt > K)[0]
        fin2, st2, cnt2, ev2 = eng.trim_many(ions[longer], 99, 0, int(cnt[longer].max()), uids=1000 + longer)
        assert np.array_equal(cnt2, cnt[longer])
        for k, i in enumerate(longer):
            f1, s1, e1 = eng.trim_one(ions[i], 99, 1000 + i, capacity=1 << 16)
            assert ev2[k, :cnt2[k]].tobytes() == e1.tobytes() and st2[k] == s1 and fin2[k].tobytes() == f1.tobytes()


@pytest.mark.parametrize("sample", ["wire", "burried_wire"])
def test_wire_samples_on_the_gpu(sample):
    """SampleWire (CUT boundaries, vacuum outside the cylinder: sample_wire.C:29-46) and SampleBurriedWire (INF
    boundaries, cover layer, matrix around the wire: sample_burried_wire.C:29-55) on the GPU (all-options kernel) against
    the oracle, whose look-ups are pinned bit for bit on the reference's records of the same cases
    (tests/golden/ref_geometry_*.npz): per-primary state (MOVING on a vacuum exit, LOST on a CUT boundary), end point,
    counters of every exit."""
    from tests.golden.make_golden import GEOMETRY_CASES
    _, box, mats, ion, start, _ = GEOMETRY_CASES[sample]
    n = 1500
    cfg = dict(tally_mask=capi.TALLY_RECORDS)
    with capi.Engine(**cfg) as eng, util.OracleEngine(util.ORC_RNG_PHILOX, **cfg) as orc:
        for e in (eng, orc):
            e.set_materials(mats)
            if sample == "wire":
                e.set_geometry(capi.GEOM_WIRE, box, bc=(capi.BC_CUT, capi.BC_CUT, capi.BC_PBC))
            else:
                e.set_geometry(capi.GEOM_BURIED_WIRE, box, bc=(capi.BC_INF, capi.BC_INF, capi.BC_INF))
        ions = capi.make_ions(n, ion[0], ion[1], ion[2], pos=start[:3], direction=start[3:])
        rg = eng.run(ions, seed=303, records=True)
        ro = orc.run(ions, seed=303, records=True)
        cg, co = eng.counters(), orc.counters()
    assert len(set(ro["state"].tolist())) >= 2
    r = util.compare_records(rg, ro, ions)
    print("%s: %s" % (sample, r))
    assert n - r["cascades_identical"] <= max_flipped_cascades(n, co["steps"]), r
    assert r["pos_outliers"] <= 0.002 * n + 1 and r["median_rel_pos"] < 0.1 * TOL, r
    for k in ("steps", "ions", "left_sample", "lost", "vacancies_created"):
        assert abs(cg[k] - co[k]) <= 2e-3 * co[k] + 2, (k, cg[k], co[k])


def test_results_do_not_depend_on_what_the_engine_ran_before():
    """GPU twin of the host test of the same name: registered primary species (class rows in shared memory) and
    unregistered ones (per-lane rows) get bit-identical constants, so a batch gives the same records on a handle that
    has already run other species as on a fresh one (what makes mytrim_uo2's files independent of the GPU count)."""
    from tests import parity_cases
    cfg = dict(tally_mask=capi.TALLY_RECORDS | capi.TALLY_PHONON)
    with capi.Engine(**cfg) as used, capi.Engine(**cfg) as fresh:
        first = parity_cases.setup_case(used, "uo2_fission_like", 400)
        parity_cases.setup_case(fresh, "uo2_fission_like", 400)
        second = parity_cases.fission_like_primaries(400, seed=11)
        used.run(first, seed=5, first_index=0)
        ra = used.run(second, seed=5, first_index=1000, records=True)
        rb = fresh.run(second, seed=5, first_index=1000, records=True)
    for f in ra.dtype.names:
        if f in ("Eel", "Enuc"):   # per-lane partial sums of a shared cascade: order of addition is not fixed
            assert np.allclose(ra[f], rb[f], rtol=1e-12, atol=0), f
        else:
            assert np.array_equal(ra[f], rb[f]), f


def test_engine_reports_the_variant_it_selects():
    """mtb_kernel_variant (C ABI) agrees with the host build of pick_variant() on every case of
    tests/test_device_loop_host.py::VARIANT_CASES."""
    from tests.test_device_loop_host import VARIANT_CASES, setup_variant_case
    for sample, cfg, want in VARIANT_CASES:
        with capi.Engine(**cfg) as eng:
            setup_variant_case(eng, sample)
            assert eng.kernel_variant() == want, (sample, cfg)
