"""Test helpers: drive the CPU oracle, the host simulator of the device loop and the compiled
reference with the same plain-struct inputs the CUDA engine takes."""
import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

from mytrim_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle.so")
HOSTSIM_LIB = os.path.join(ROOT, "tests", "libhostsim.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_UO2 = os.path.join(REF_DIR, "mytrim_uo2")
GOLDEN = os.path.join(ROOT, "tests", "golden")

ORC_RNG_MT19937, ORC_RNG_PHILOX = 0, 1

_libs = {}


def _load(path, prefix):
    if path not in _libs:
        lib = C.CDLL(path)
        capi.declare(lib, prefix)
        _libs[path] = lib
    return _libs[path]


class OracleEngine(capi.EngineBase):
    _prefix = "orc_"

    def __init__(self, rng_mode=ORC_RNG_PHILOX, config=None, **kw):
        lib = _load(ORACLE_LIB, "orc_")
        cfg = config if config is not None else capi.default_config(**kw)
        lib.orc_create.argtypes = [C.POINTER(capi.Config), C.c_int]
        lib.orc_create.restype = C.c_void_p
        lib.orc_getrstop.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        lib.orc_getrstop.restype = C.c_double
        lib.orc_average.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_size_t]
        lib.orc_average.restype = C.c_int
        lib.orc_destroy.restype = None
        super().__init__(lib, C.c_void_p(lib.orc_create(C.byref(cfg), rng_mode)))
        self.config = cfg

    def stopping(self, material, Z1, m1, E):
        Z1 = np.broadcast_to(np.asarray(Z1), np.shape(E))
        m1 = np.broadcast_to(np.asarray(m1, dtype=float), np.shape(E))
        return np.array([self._lib.orc_getrstop(self._h, material, int(z), float(m), float(e))
                         for z, m, e in zip(Z1, m1, np.asarray(E, dtype=float))])

    def average(self, material, Z1, m1, n_elements):
        out = np.zeros(6 + 4 * n_elements)
        self._check(self._lib.orc_average(self._h, material, Z1, m1, out.ctypes.data, len(out)))
        return out

    def range_list(self, capacity=1 << 22):
        x = np.zeros(capacity, dtype=np.float64)
        z = np.zeros(capacity, dtype=np.int32)
        n = C.c_size_t()
        self._check(self._lib.orc_get_range_list(self._h, x.ctypes.data, z.ctypes.data, capacity, C.byref(n)))
        return x[:n.value].copy(), z[:n.value].copy()


class HostSimEngine(capi.EngineBase):
    """The device lane loop compiled for the host (tests/hostsim.cpp)."""
    _prefix = "hs_"

    def __init__(self, config=None, **kw):
        lib = _load(HOSTSIM_LIB, "hs_")
        cfg = config if config is not None else capi.default_config(**kw)
        lib.hs_create.argtypes = [C.POINTER(capi.Config)]
        lib.hs_create.restype = C.c_void_p
        lib.hs_error.argtypes = [C.c_void_p]
        lib.hs_error.restype = C.c_char_p
        lib.hs_destroy.restype = None
        super().__init__(lib, C.c_void_p(lib.hs_create(C.byref(cfg))))
        self.config = cfg

    def _error_text(self):
        return self._lib.hs_error(self._h).decode()

    def reset_tallies(self):
        raise NotImplementedError

    def stopping(self, material, Z1, m1, E):
        E = np.ascontiguousarray(E, dtype=np.float64)
        Z1 = np.ascontiguousarray(np.broadcast_to(Z1, E.shape), dtype=np.int32)
        m1 = np.ascontiguousarray(np.broadcast_to(m1, E.shape), dtype=np.float64)
        out = np.zeros(len(E))
        self._check(self._lib.hs_stopping(self._h, material, len(E), Z1.ctypes.data, m1.ctypes.data,
                                          E.ctypes.data, out.ctypes.data))
        return out

    def range_list(self, capacity=1 << 22):
        x = np.zeros(capacity, dtype=np.float32)
        z = np.zeros(capacity, dtype=np.int32)
        n = C.c_size_t()
        self._check(self._lib.hs_get_range_list(self._h, x.ctypes.data, z.ctypes.data, capacity, C.byref(n)))
        return x[:n.value].copy(), z[:n.value].copy()


def have_reference():
    return os.path.exists(REF_DRIVER)


def ref_env():
    env = dict(os.environ)
    env["MYTRIM_DATADIR"] = os.path.join(REF_DIR, "data")
    return env


def run_reference(script, timeout=600):
    """Feeds a command script to oracle/_ref/ref_driver; returns stdout lines."""
    p = subprocess.run([REF_DRIVER], input=script, capture_output=True, text=True, env=ref_env(), timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("ref_driver failed: " + p.stderr)
    return p.stdout.strip().split("\n")


def reference_script(ion, materials, thicknesses, n=0, tally="vaccount", threads=1, seeds_file=None, master=2344,
                     out=None, primaries_only=False, box=None, start=None, scale=None, potential=None, sample=None, tmin=None, cw=None):
    Z, m, E = ion[:3]
    lines = ["ion %d %.17g %.17g" % (Z, m, E) + (" %.17g" % ion[3] if len(ion) > 3 else ""), "n %d" % n, "threads %d" % threads, "tally %s" % tally,
             "master %d" % master, "primaries_only %d" % int(primaries_only)]
    if scale is not None:
        lines.append("scale %.17g" % scale)
    if potential is not None:
        lines.append("potential %s" % potential)
    if sample is not None:
        lines.append("sample %s" % sample)
    if tmin is not None:
        lines.append("tmin %.17g" % tmin)
    if cw is not None:
        lines.append("cw %.17g" % cw)
    if seeds_file:
        lines.append("seeds %s" % seeds_file)
    if box is not None:
        lines.append("box %.17g %.17g %.17g" % tuple(box))
    if start is not None:
        lines.append("start " + " ".join("%.17g" % v for v in start))
    if out:
        lines.append("out %s" % out)
    for mat, th in zip(materials, thicknesses):
        lines.append("layer %.17g %.17g %d" % (th, mat["rho"], len(mat["elements"])))
        for e in mat["elements"]:
            lines.append("elem %d %.17g %.17g %.17g %.17g" % (e["Z"], e["m"], e["t"], e.get("Edisp", 25.0),
                                                            e.get("Elbind", 3.0)))
    return lines


def run_reference_cascades(ion, materials, thicknesses, seeds, tally="vaccount", threads=1, timeout=600, **kw):
    """Runs the unmodified reference for the given per-primary seeds; returns (records, summary dict)."""
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    with tempfile.TemporaryDirectory() as tmp:
        sf = os.path.join(tmp, "seeds.bin")
        seeds.tofile(sf)
        out = os.path.join(tmp, "out")
        lines = reference_script(ion, materials, thicknesses, n=len(seeds), tally=tally, threads=threads,
                                 seeds_file=sf, out=out, **kw)
        lines.append("run")
        stdout = run_reference("\n".join(lines) + "\n", timeout=timeout)
        rec = np.fromfile(out + ".records", dtype=capi.RECORD_DTYPE)
        hist = None
        if os.path.exists(out + "_vac.dat"):
            hist = np.loadtxt(out + "_vac.dat", ndmin=2)
        elif os.path.exists(out + "_evac.dat"):   # TrimVacEnergyCount::writeOutput: "E x count" lines
            hist = np.loadtxt(out + "_evac.dat", ndmin=2)
        elif os.path.exists(out + "_ranges.dat"):  # TrimRange::writeOutput: "#x Z.." header, then "x count.." lines
            with open(out + "_ranges.dat") as f:
                hist = f.read()
    return rec, json.loads(stdout[-1]), hist


def distinct_seeds(n, master=2344):
    """32-bit distinct per-primary seeds (SURVEY.md §8c caveat: runmytrim's irand() seeds are 16-bit)."""
    x = (np.arange(n, dtype=np.uint64) + np.uint64((master * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return (x & np.uint64(0xFFFFFFFF)).astype(np.uint32)


# --- the BASELINE.json configurations: mytrim_b200/workloads.py ---------------------------
from mytrim_b200.workloads import (CONFIGS, CU, FE, UO2, W, XE_GAS, ZRO2, primaries_for,  # noqa: E402,F401
                                   setup_engine)


# --- north-star statistical criterion against a SUMMARY of a large reference sample -------------------------
# 1e6 per-primary records of the reference are 72 MB; the committed fixture keeps, per observable, either
# K quantiles (continuous observables) or the exact histogram (integer observables), plus mean and variance.
STAT_K = 4096
STAT_CONTINUOUS = {"projected_range_x": lambda r: r["pos"][:, 0],
                   "lateral_range": lambda r: np.hypot(r["pos"][:, 1] - 50.0, r["pos"][:, 2] - 50.0),
                   "electronic_loss_Eel": lambda r: r["Eel"]}
STAT_INTEGER = {"vacancies_per_ion": lambda r: r["vacancies"], "replacements_per_ion": lambda r: r["replacements"],
                "collision_steps": lambda r: r["steps"], "ions_followed": lambda r: r["ions"]}


def summarize_records(rec):
    """Summary of per-primary records for ks_against_summary()."""
    out = {"n": np.array(len(rec))}
    probs = (np.arange(STAT_K) + 0.5) / STAT_K
    for name, get in STAT_CONTINUOUS.items():
        v = np.asarray(get(rec), dtype=np.float64)
        out["q_" + name] = np.quantile(v, probs)
        out["m_" + name] = np.array([v.mean(), v.var()])
    for name, get in STAT_INTEGER.items():
        v = np.asarray(get(rec), dtype=np.int64)
        out["h_" + name] = np.bincount(v)
        out["m_" + name] = np.array([v.mean(), v.var()])
    return out


def ks_against_summary(rec, summary):
    """{observable: (mean, reference mean, KS D, asymptotic two-sample p)} of records against a summary."""
    from scipy import stats
    m = int(summary["n"])
    n = len(rec)
    res = {}
    scale = np.sqrt(n * m / float(n + m))
    probs = (np.arange(STAT_K) + 0.5) / STAT_K
    for name, get in STAT_CONTINUOUS.items():
        v = np.sort(np.asarray(get(rec), dtype=np.float64))
        F = np.searchsorted(v, summary["q_" + name], side="right") / float(n)
        D = float(np.abs(F - probs).max())   # resolution 1/(2K) = 1.2e-4, 5 % of the critical D at 1e6 ions
        res[name] = (v.mean(), float(summary["m_" + name][0]), D, float(stats.distributions.kstwobign.sf(D * scale)))
    for name, get in STAT_INTEGER.items():
        h = np.bincount(np.asarray(get(rec), dtype=np.int64))
        g = np.asarray(summary["h_" + name], dtype=np.float64)
        L = max(len(h), len(g))
        ch = np.cumsum(np.pad(h, (0, L - len(h)))) / float(n)
        cg = np.cumsum(np.pad(g, (0, L - len(g)))) / float(m)
        D = float(np.abs(ch - cg).max())
        mean = float((np.arange(len(h)) * h).sum() / n)
        res[name] = (mean, float(summary["m_" + name][0]), D, float(stats.distributions.kstwobign.sf(D * scale)))
    return res


# --- deterministic criterion, per ION (north_star: "per-ion trajectories match, within 1e-5 relative FP tolerance, a
# CPU replay that uses the same Philox stream") -------------------------------------------------------------------
def compare_ion_logs(a, b, primaries, first_index=0, tol=1e-5):
    """Joins two ion logs (MTB_TALLY_IONLOG: one entry per followed ion with its birth and death state) by the
    scheduling-independent ion id and compares EVERY ion: integer fields (primary, Z, generation, tag, final state)
    bit for bit; birth and death position relative to the distance from the primary's source point (the scale of
    the coordinates the FP32 arithmetic carried); birth and final energy relative to the primary's energy and,
    separately, to the ion's own energy.  `a` is the path under test, `b` the replay.  Returns the counts."""
    assert len(np.unique(a["uid"])) == len(a) and len(np.unique(b["uid"])) == len(b), "ion ids collide"
    common, xa, xb = np.intersect1d(a["uid"], b["uid"], return_indices=True)
    A, B = a[xa], b[xb]
    ints = np.ones(len(A), dtype=bool)
    for f in ("primary", "Z", "gen", "tag", "state"):
        ints &= A[f] == B[f]
    src = primaries[(B["primary"] - first_index).astype(np.int64)]
    scale = np.maximum(np.linalg.norm(B["pos1"] - src["pos"], axis=1), 1.0)
    d0 = np.linalg.norm(A["pos0"] - B["pos0"], axis=1) / scale
    d1 = np.linalg.norm(A["pos1"] - B["pos1"], axis=1) / scale
    e0 = np.abs(A["E0"] - B["E0"])
    e1 = np.abs(A["E1"] - B["E1"])
    pos_ok = (d0 < tol) & (d1 < tol)
    e_prim_ok = (e0 < tol * src["E"]) & (e1 < tol * src["E"])
    e_own_ok = e0 <= tol * B["E0"]
    return dict(n_test=len(a), n_replay=len(b), joined=len(common), ints_equal=int(ints.sum()),
                pos_outliers=int((ints & ~pos_ok).sum()), energy_outliers=int((ints & ~e_prim_ok).sum()),
                own_energy_outliers=int((ints & ~e_own_ok).sum()),
                all_ok=int((ints & pos_ok & e_prim_ok).sum()),
                median_rel_pos=float(np.median(d1[ints])) if ints.any() else 0.0,
                max_rel_pos=float(np.maximum(d0, d1)[ints].max()) if ints.any() else 0.0)


def compare_records(ra, rb, primaries, tol=1e-5):
    """Per-primary records of two runs: cascades whose integer fields all agree, and among the primaries with the
    same number of collisions the ones whose end point differs by more than tol of the distance travelled."""
    ints = np.ones(len(ra), dtype=bool)
    for f in ("vacancies", "replacements", "steps", "ions", "state", "primary_steps"):
        ints &= ra[f] == rb[f]
    sel = ra["primary_steps"] == rb["primary_steps"]
    scale = np.maximum(np.linalg.norm(rb["pos"] - primaries["pos"], axis=1), 1.0)
    rel = np.linalg.norm(ra["pos"] - rb["pos"], axis=1) / scale
    eel = np.abs(ra["Eel"] - rb["Eel"]) > tol * np.maximum(rb["Eel"], 1e-300)
    return dict(n=len(ra), cascades_identical=int(ints.sum()), same_primary_steps=int(sel.sum()),
                pos_outliers=int((sel & (rel >= tol)).sum()), eel_outliers=int((ints & eel).sum()),
                median_rel_pos=float(np.median(rel[sel])) if sel.any() else 0.0)
