#!/usr/bin/env python3
"""Generates the committed golden fixtures from the UNMODIFIED reference compiled in this
container (oracle/_ref, see oracle/Makefile).  Run from the repo root:

    python tests/golden/make_golden.py

Fixtures:
  uo2_out.{Erec,clcoor,dist}  `MYTRIM_SEED=39172 mytrim_uo2 out 10 0.1 1` (tests/uo2/test.sh);
                               byte-identical to the reference's tests/uo2/gold/ files.
  uo2_out2.{Erec,clcoor,dist} `MYTRIM_SEED=4711 mytrim_uo2 out2 6 0.5 2` (22 bubbles, two fission events).
  rng_mt19937.txt             SimconfType::drand()/irand() sequences (simconf.h:52-53).
  stopping.json               MaterialBase::getrstop + average() known answers.
  ref_records_<cfg>.npz       per-primary records of the reference for fixed 32-bit seeds.
  ref_records_<cfg>_<potential>.npz  the same with TrimBase::_potential = MOLIERE / CKR.
  ref_tally_<tally>_<cfg>.npz records (+ output file content) of TrimRange / TrimPrimaries / TrimRecoils /
                               TrimVacEnergyCount / TrimPhononOut.
  ref_geometry_<sample>.npz    records of the reference in a SampleWire / SampleBurriedWire.
  ref_options_scale10.npz      records with a length scale of 10 A, per-element Edisp / Elbind and Ef = 5 eV.
  ref_records_layer_stack.npz  Cu / Fe / W / ZrO2 stack: the layer look-up with different materials.
  ref_stats_<cfg>.npz          quantiles / histograms / means of 1e5..1e6 reference cascades per configuration
                               (statistical criterion; STATISTICS_CASES).
  ref_options_tmin1_cw0p01.npz, ref_options_primaries_only.npz   tmin = 1, cw = 0.01; ThreadedTrimBase::_primaries_only.
  ref_output_evac_<cfg>.npz, ref_output_ranges_<cfg>.npz   content of the files the reference's writeOutput() wrote for
                               validation/c_on_w/input.json (2000 primaries) and validation/cu_on_cu/cu_on_cu.json (1e5).
  ref_uo2_seed777.npz          summary of .Erec / .dist of `MYTRIM_SEED=777 mytrim_uo2 out 10 1.0 100` (44 bubbles).
  vacancy_count_published.json the reference's published vacancies/ion table
                               (validation/vacancy_count/vacancy_count_comparison.dat).
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def uo2():
    with tempfile.TemporaryDirectory() as tmp:
        env = util.ref_env()
        env["MYTRIM_SEED"] = "39172"
        subprocess.run([util.REF_UO2, "out", "10", "0.1", "1"], cwd=tmp, env=env, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for ext in ("Erec", "clcoor", "dist"):
            shutil.copy(os.path.join(tmp, "out." + ext), os.path.join(HERE, "uo2_out." + ext))
        # a second experiment of the same app: smaller, denser bubbles (22 clusters), two fission events
        env["MYTRIM_SEED"] = "4711"
        subprocess.run([util.REF_UO2, "out2", "6", "0.5", "2"], cwd=tmp, env=env, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for ext in ("Erec", "clcoor", "dist"):
            shutil.copy(os.path.join(tmp, "out2." + ext), os.path.join(HERE, "uo2_out2." + ext))


def rng():
    lines = util.run_reference("rng 39172 64\nrng 2344 64\n")
    with open(os.path.join(HERE, "rng_mt19937.txt"), "w") as f:
        f.write("# seed 39172: 64 drand (hexfloat) then 64 irand; then the same for seed 2344\n")
        f.write("\n".join(lines) + "\n")


STOPPING_CASES = [
    # (name, ion Z, m, material, energies)
    ("cu_on_cu", 29, 63.546, util.CU, [25.0, 1e2, 1e3, 1e4, 1.5e5, 1e6, 1e7, 1e8]),
    ("h_on_fe", 1, 1.008, util.FE, [1e2, 1e3, 1e4, 2e4, 3e4, 1e5, 1e6, 1e7]),
    ("he_on_fe", 2, 4.003, util.FE, [1e2, 3e3, 4e3, 5e3, 1e4, 1e5, 1e6, 1e7]),
    ("c_on_w", 6, 12.0, util.W, [1e2, 1e4, 1e5, 1e6, 1e7, 1e8]),
    ("xe_on_uo2", 54, 131.904, {"rho": 10.97, "elements": [{"Z": 92, "m": 238.03, "t": 1}, {"Z": 8, "m": 15.999, "t": 2}]},
     [1e3, 1e5, 1e7, 1e8, 1e9]),
    ("si_on_c", 14, 28.086, {"rho": 2.26, "elements": [{"Z": 6, "m": 12.011, "t": 1}]}, [1e2, 1e4, 1e6, 1e8]),
    ("o_on_zro2", 8, 16.0, util.ZRO2, [30.0, 1e3, 1e5, 1e7]),
    ("u_on_uo2", 92, 235.0, util.UO2, [50.0, 1e4, 1e6, 1e8]),
]


def stopping():
    out = {}
    for name, Z, m, mat, energies in STOPPING_CASES:
        lines = util.reference_script((Z, m, 0.0), [mat], [1000.0])
        for E in energies:
            lines.append("stopping %d %.17g %.17g" % (Z, m, E))
        lines.append("average %d %.17g" % (Z, m))
        res = util.run_reference("\n".join(lines) + "\n")
        vals = [float(l.split()[4]) for l in res if l.startswith("stopping")]
        avg = [float(x) for x in [l for l in res if l.startswith("average")][0].split()[3:]]
        out[name] = {"Z": Z, "m": m, "material": mat, "E": energies, "getrstop": vals, "average": avg}
    with open(os.path.join(HERE, "stopping.json"), "w") as f:
        json.dump(out, f, indent=1)


RECORD_CASES = {"cu_on_cu_10keV": 256, "cu_on_cu_1keV": 512, "h_on_fe_100keV": 256, "he_on_fe_100keV": 96,
                "c_on_w_1MeV": 24, "xe_on_zro2_500keV": 12}


def records():
    for name, n in RECORD_CASES.items():
        c = util.CONFIGS[name]
        seeds = util.distinct_seeds(n)
        rec, summary, hist = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], seeds,
                                                         box=c.get("box"))
        np.savez_compressed(os.path.join(HERE, "ref_records_%s.npz" % name), records=rec, seeds=seeds,
                            vac=hist[:, 1].astype(np.uint64), repl=hist[:, 2].astype(np.uint64),
                            summary=json.dumps(summary))


# the other two interatomic potentials of trim.C:194-222, 238-259 (TrimBase::_potential)
POTENTIAL_CASES = {"moliere": ("cu_on_cu_1keV", 256), "ckr": ("cu_on_cu_1keV", 256)}


def potentials():
    for pot, (name, n) in POTENTIAL_CASES.items():
        c = util.CONFIGS[name]
        seeds = util.distinct_seeds(n, master=77)
        rec, summary, hist = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], seeds,
                                                         box=c.get("box"), potential=pot)
        np.savez_compressed(os.path.join(HERE, "ref_records_%s_%s.npz" % (name, pot)), records=rec, seeds=seeds,
                            vac=hist[:, 1].astype(np.uint64), repl=hist[:, 2].astype(np.uint64),
                            summary=json.dumps(summary))


# the other in-tree tally classes: (reference tally name, configuration, primaries)
TALLY_CASES = {"range": ("cu_on_cu_1keV", 192), "primaries": ("cu_on_cu_10keV", 64), "recoils": ("cu_on_cu_10keV", 64),
               "vacenergycount": ("c_on_w_1MeV", 12), "phonon": ("cu_on_cu_1keV", 128)}
# apps/mytrim_layers.C: TrimRecoils on the 50-layer ZrO2 stack of inputs/samplelayers_zro2_multilayer.in
TALLY_CASES_EXTRA = {"recoils": ("xe_on_zro2_500keV", 8)}


def tallies():
    for tally, (name, n) in list(TALLY_CASES.items()) + list(TALLY_CASES_EXTRA.items()):
        c = util.CONFIGS[name]
        seeds = util.distinct_seeds(n, master=101)
        rec, summary, hist = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], seeds,
                                                         box=c.get("box"), tally=tally)
        extra = {}
        if tally == "vacenergycount":
            extra["evac"] = hist[hist[:, 2] > 0].astype(np.int64)   # non-zero rows (E bin, x bin, count)
        if tally == "range":
            extra["ranges_dat"] = np.array(hist)       # the text of <base>_ranges.dat
        np.savez_compressed(os.path.join(HERE, "ref_tally_%s_%s.npz" % (tally, name)), records=rec, seeds=seeds,
                            summary=json.dumps(summary), **extra)


# SampleWire (CUT boundaries in x, y; vacuum outside the cylinder) and SampleBurriedWire (INF boundaries, cover
# layer, matrix around the wire): (sample class, box, materials [wire, cover/matrix], ion, start, primaries)
GEOMETRY_CASES = {
    "wire": ("wire", (60.0, 60.0, 1000.0), [util.CU], (29, 63.546, 2.0e4), (30.0, 18.0, 0.0, 0.0, 0.3, 1.0), 192),
    "burried_wire": ("burried_wire", (100.0, 100.0, 300.0), [util.CU, util.FE], (29, 63.546, 2.0e4),
                     (50.0, 50.0, -200.0, 0.1, 0.0, 1.0), 128),
}


def geometries():
    for key, (sample, box, mats, ion, start, n) in GEOMETRY_CASES.items():
        seeds = util.distinct_seeds(n, master=303)
        rec, summary, _ = util.run_reference_cascades(ion, mats, [1.0] * len(mats), seeds, box=box, sample=sample,
                                                      start=start)
        np.savez_compressed(os.path.join(HERE, "ref_geometry_%s.npz" % key), records=rec, seeds=seeds,
                            summary=json.dumps(summary))


# run-time options of the path: SimconfType::setLengthScale (positions in units of 10 A), per-element Edisp / Elbind
# (runmytrim.C:246-250) and the final energy IonBase::_Ef
OPTION_MATERIAL = {"rho": 8.92, "elements": [{"Z": 29, "m": 63.546, "t": 1.0, "Edisp": 30.0, "Elbind": 2.0}]}
OPTION_CASE = dict(ion=(29, 63.546, 2.0e4, 5.0), scale=10.0, box=(100.0, 10.0, 10.0), n=96)


def options():
    o = OPTION_CASE
    seeds = util.distinct_seeds(o["n"], master=404)
    rec, summary, hist = util.run_reference_cascades(o["ion"], [OPTION_MATERIAL], [o["box"][0]], seeds, box=o["box"],
                                                     scale=o["scale"])
    np.savez_compressed(os.path.join(HERE, "ref_options_scale10.npz"), records=rec, seeds=seeds,
                        vac=hist[:, 1].astype(np.uint64), repl=hist[:, 2].astype(np.uint64), summary=json.dumps(summary))


def options2():
    # SimconfType::tmin / cw (simconf.C:45-47, trim.C:88-94) and ThreadedTrimBase::_primaries_only
    c = util.CONFIGS["cu_on_cu_10keV"]
    for key, kw in (("tmin1_cw0p01", dict(tmin=1.0, cw=0.01)), ("primaries_only", dict(primaries_only=True))):
        seeds = util.distinct_seeds(96, master=505)
        rec, summary, hist = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"], seeds, **kw)
        np.savez_compressed(os.path.join(HERE, "ref_options_%s.npz" % key), records=rec, seeds=seeds,
                            vac=hist[:, 1].astype(np.uint64), repl=hist[:, 2].astype(np.uint64), summary=json.dumps(summary))


# a stack of DIFFERENT materials (sample_layers.C:26-49: linear scan, x < 0 -> first layer, beyond -> last layer)
STACK_CASE = dict(ion=(29, 63.546, 2.0e4), materials=[util.CU, util.FE, util.W, util.ZRO2],
                  thicknesses=[30.0, 40.0, 60.0, 500.0], n=128)


def stack():
    o = STACK_CASE
    seeds = util.distinct_seeds(o["n"], master=606)
    rec, summary, hist = util.run_reference_cascades(o["ion"], o["materials"], o["thicknesses"], seeds)
    np.savez_compressed(os.path.join(HERE, "ref_records_layer_stack.npz"), records=rec, seeds=seeds,
                        vac=hist[:, 1].astype(np.uint64), repl=hist[:, 2].astype(np.uint64), summary=json.dumps(summary))


# samples of the north-star statistical criterion: cascades of the unmodified reference with distinct 32-bit seeds
# (SURVEY.md §8c), summarised (tests/util.py::summarize_records).  1e6 Cu->Cu 10 keV cascades take ~4 min on 8 cores.
STATISTICS_CASES = {"cu_on_cu_10keV": 1000000, "cu_on_cu_1keV": 1000000, "h_on_fe_100keV": 500000,
                    "he_on_fe_100keV": 100000, "c_on_w_1MeV": 100000, "xe_on_zro2_500keV": 100000,
                    # the file-energy / long-cascade configurations (tests/json/cu_on_cu.json, H->Fe at 1 MeV,
                    # tests/json/xe_on_uo2.json)
                    "cu_on_cu_150keV": 100000, "h_on_fe_1MeV": 200000, "xe_on_uo2_10MeV": 10000}


def statistics(only=None):
    for name, n in STATISTICS_CASES.items():
        if only and name not in only:
            continue
        c = util.CONFIGS[name]
        rec, summary, _ = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"],
                                                      util.distinct_seeds(n, master=1), threads=os.cpu_count() or 1,
                                                      box=c.get("box"), timeout=7200)
        np.savez_compressed(os.path.join(HERE, "ref_stats_%s.npz" % name), summary=json.dumps(summary),
                            **util.summarize_records(rec))


# what the reference's own writeOutput() puts into <base>_evac.dat (TrimVacEnergyCount.C:70-81) and <base>_ranges.dat
# (TrimRange.C:66-121) for samples large enough to compare the GPU drivers' files with them statistically
# (SURVEY.md §8f row 1): validation/c_on_w/input.json (vacenergycount) and validation/cu_on_cu/cu_on_cu.json (range)
OUTPUT_CASES = {"evac": ("c_on_w_1MeV", "vacenergycount", 2000), "ranges": ("cu_on_cu_150keV", "range", 100000)}


def outputs():
    for key, (name, tally, n) in OUTPUT_CASES.items():
        c = util.CONFIGS[name]
        rec, summary, hist = util.run_reference_cascades(c["ion"], c["materials"], c["thicknesses"],
                                                         util.distinct_seeds(n, master=707), tally=tally,
                                                         threads=os.cpu_count() or 1, timeout=3600)
        if key == "evac":
            np.savez_compressed(os.path.join(HERE, "ref_output_evac_%s.npz" % name), n=n,
                                evac=hist[hist[:, 2] > 0].astype(np.int64), shape=hist[:, :2].max(axis=0).astype(np.int64) + 1,
                                vacancies=int(rec["vacancies"].sum()))
        else:
            # <base>_ranges.dat of 1e5 primaries has 2e5 lines: keep its header, bin geometry, column totals and the
            # quantiles of every column's distribution
            rows = [l.split() for l in str(hist).strip().split("\n")]
            table = np.array([[float(v) for v in r] for r in rows[1:]])
            probs = (np.arange(2048) + 0.5) / 2048
            cdf = np.cumsum(table[:, 1:], axis=0) / table[:, 1:].sum(axis=0)
            q = np.array([table[np.searchsorted(cdf[:, k], probs), 0] for k in range(cdf.shape[1])])
            np.savez_compressed(os.path.join(HERE, "ref_output_ranges_%s.npz" % name), n=n, header=" ".join(rows[0]),
                                nbin=len(table), x_min=table[0, 0], bin_width=table[1, 0] - table[0, 0],
                                totals=table[:, 1:].sum(axis=0), quantiles=q)


def published():
    src = "/root/reference/validation/vacancy_count/vacancy_count_comparison.dat"
    rows = [l.strip().split(",") for l in open(src)][2:]
    rows = [r for r in rows if len(r) >= 13 and r[0]]
    data = {"energy_keV": [float(r[0]) for r in rows],
            "si_on_c_kp": [float(r[1]) for r in rows], "si_on_c_exact": [float(r[2]) for r in rows],
            "xe_on_u_kp": [float(r[5]) for r in rows], "xe_on_u_exact": [float(r[6]) for r in rows],
            "cu_on_cu_kp": [float(r[9]) for r in rows], "cu_on_cu_exact": [float(r[10]) for r in rows],
            "source": "validation/vacancy_count/vacancy_count_comparison.dat (MyTRIM: KP / MyTRIM: exact columns)"}
    with open(os.path.join(HERE, "vacancy_count_published.json"), "w") as f:
        json.dump(data, f, indent=1)


if __name__ == "__main__":
    import __graft_entry__ as g
    g.build_test_infrastructure()
    uo2()
    rng()
    stopping()
    records()
    potentials()
    tallies()
    geometries()
    options()
    options2()
    stack()
    statistics()
    outputs()
    published()
    print("golden fixtures written to", HERE)
