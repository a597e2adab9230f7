"""ctypes mirror of include/mytrim_b200.h.

Python is only the test/bench harness language here: the product is the C-ABI shared library
``libmytrim_b200.so`` (CUDA kernels + extern "C" layer) and the C++ façade in ``include/mytrim``.
The structures below follow the header field by field; ``Engine`` is a thin convenience wrapper
that keeps the call sequence of the reference apps (build materials -> build sample -> run
primaries -> read tallies, apps/runmytrim.C:183-331).

The same struct definitions are reused by the test helpers (tests/util.py) to feed the CPU
checker identical inputs; nothing in this package loads or calls it.
"""
import ctypes as C
import os

import numpy as np

NZ = 92
VMAP_NX = 20
VMAP_NY = 20

OK, EINVAL, ECUDA, ENODEV, ESTACK, ENOMEM, ECAPACITY, ENCCL = range(8)

POT_UNIVERSAL, POT_MOLIERE, POT_CKR = 0, 1, 2
BC_PBC, BC_INF, BC_CUT = 0, 1, 2
GEOM_SOLID, GEOM_LAYERS, GEOM_WIRE, GEOM_BURIED_WIRE, GEOM_CLUSTERS = range(5)
MOVING, REPLACEMENT, SUBSTITUTIONAL, INTERSTITIAL, LOST, DELETE, VACANCY = range(7)
FOLLOW_ALL, FOLLOW_NONE, FOLLOW_GEN_LT = 0, 1, 2
VAC_COUNT, VAC_NRT, VAC_KP, VAC_NONE = 0, 1, 2, 3
TALLY_VAC_DEPTH = 1 << 0
TALLY_VAC_ENERGY = 1 << 1
TALLY_RANGE = 1 << 2
TALLY_PHONON = 1 << 3
TALLY_VACMAP = 1 << 4
TALLY_RECORDS = 1 << 5
TALLY_IONLOG = 1 << 6


class Config(C.Structure):
    _fields_ = [
        ("tmin", C.c_double), ("tau", C.c_double), ("cw", C.c_double), ("length_scale", C.c_double),
        ("potential", C.c_int32), ("follow", C.c_int32), ("follow_max_gen", C.c_int32),
        ("vacancy_model", C.c_int32), ("tally_mask", C.c_uint32), ("vmap_z", C.c_int32 * 3),
        ("ionlog_z", C.c_int32), ("hist_bins", C.c_int32), ("evac_rows", C.c_int32), ("device", C.c_int32),
        ("ionlog_capacity", C.c_uint64), ("range_capacity", C.c_uint64),
    ]


class Element(C.Structure):
    _fields_ = [("Z", C.c_int32), ("_pad", C.c_int32), ("m", C.c_double), ("t", C.c_double),
                ("Edisp", C.c_double), ("Elbind", C.c_double)]


class Material(C.Structure):
    _fields_ = [("rho", C.c_double), ("tag", C.c_int32), ("n_elements", C.c_int32),
                ("first_element", C.c_int32), ("_pad", C.c_int32)]


class Geometry(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("bc", C.c_int32 * 3), ("w", C.c_double * 3),
        ("n_layers", C.c_int32), ("_pad0", C.c_int32), ("layer_thickness", C.POINTER(C.c_double)),
        ("kn", C.c_int32 * 3), ("n_clusters", C.c_int32), ("cluster_xyzr", C.POINTER(C.c_double)),
    ]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "vacancies_created", "replacements", "steps", "ions", "primaries", "recoils_queued", "lost",
        "left_sample", "hist_clamped", "stack_max")] + [("EelTotal", C.c_double), ("EnucTotal", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


ION_DTYPE = np.dtype([
    ("pos", "f8", 3), ("dir", "f8", 3), ("E", "f8"), ("m", "f8"), ("Ef", "f8"),
    ("Z", "i4"), ("gen", "i4"), ("tag", "i4"), ("seed", "u4")], align=True)
RECORD_DTYPE = np.dtype([
    ("pos", "f8", 3), ("E", "f8"), ("Eel", "f8"), ("Enuc", "f8"), ("vacancies", "u4"),
    ("replacements", "u4"), ("steps", "u4"), ("ions", "u4"), ("state", "i4"), ("primary_steps", "u4")], align=True)
IONLOG_DTYPE = np.dtype([
    ("pos0", "f8", 3), ("pos1", "f8", 3), ("E0", "f8"), ("E1", "f8"), ("uid", "u8"), ("primary", "u8"),
    ("Z", "i4"), ("gen", "i4"), ("tag", "i4"), ("state", "i4")], align=True)
EVENT_DTYPE = np.dtype([
    ("pka_pos", "f8", 3), ("pka_dir", "f8", 3), ("pka_E", "f8"), ("recoil_pos", "f8", 3),
    ("recoil_dir", "f8", 3), ("recoil_E", "f8"), ("ls", "f8"), ("dee", "f8"), ("den", "f8"),
    ("material", "i4"), ("element", "i4"), ("material_tag", "i4"), ("pka_state", "i4"),
    ("recoil_above_threshold", "i4"), ("_pad", "i4")], align=True)
assert ION_DTYPE.itemsize == 88 and RECORD_DTYPE.itemsize == 72
assert IONLOG_DTYPE.itemsize == 96 and EVENT_DTYPE.itemsize == 160


class MytrimError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mytrim_b200 status %d: %s" % (code, msg))
        self.code = code


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MYTRIM_B200_LIB", os.path.join(_PKG_DIR, "libmytrim_b200.so"))
_lib = None


def load_library():
    """Loads the CUDA engine.  There is deliberately no fallback of any kind."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        if "MYTRIM_B200_NCCL_LIB" not in os.environ:
            # mtb_allreduce dlopens NCCL: in a Python process it has to be the one torch is built against
            try:
                import importlib.util
                spec = importlib.util.find_spec("nvidia.nccl")
                for d in (spec.submodule_search_locations if spec else []):
                    cand = os.path.join(d, "lib", "libnccl.so.2")
                    if os.path.exists(cand):
                        os.environ["MYTRIM_B200_NCCL_LIB"] = cand
                        break
            except (ImportError, ValueError, AttributeError):
                pass
        _lib = C.CDLL(LIB_PATH)  # RTLD_LOCAL: its C++ symbols must not interpose other libraries
        declare(_lib, "mtb_")
        _lib.mtb_version.restype = C.c_char_p
        _lib.mtb_last_error.restype = C.c_char_p
    return _lib


def declare(lib, prefix):
    """Sets argtypes for the entry points of a library exporting `<prefix>run`, `<prefix>get_counters`, ..."""
    vp = C.c_void_p
    u64p = C.POINTER(C.c_uint64)
    szp = C.POINTER(C.c_size_t)

    def f(name, *args):
        fn = getattr(lib, prefix + name, None)
        if fn is not None:
            fn.argtypes = list(args)
            fn.restype = C.c_int
        return fn

    f("set_materials", vp, C.c_int, C.POINTER(Material), C.c_int, C.POINTER(Element))
    f("set_geometry", vp, C.POINTER(Geometry))
    f("set_tables", vp, vp, vp, vp, vp)
    f("run", vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, vp)
    f("run_beam", vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, vp)
    f("upload_primaries", vp, C.c_uint64, vp)
    f("launch_resident", vp, C.c_uint64, C.c_uint64)
    f("synchronize", vp)
    f("last_kernel_ms", vp, C.POINTER(C.c_float))
    f("fetch_records", vp, C.c_uint64, vp)
    f("reset_tallies", vp)
    f("get_counters", vp, C.POINTER(Counters))
    f("get_vac_depth", vp, vp, vp, C.c_size_t, szp)
    f("get_vac_energy", vp, vp, C.c_size_t, C.c_size_t)
    f("get_vacmap", vp, vp)
    f("get_range_list", vp, vp, vp, C.c_size_t, szp)
    f("get_ion_log", vp, vp, C.c_size_t, szp)
    f("hist_bins", vp, szp, szp)
    f("tally_device_views", vp, C.POINTER(vp), szp, C.POINTER(vp), szp)
    f("trim_one", vp, vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_int32), vp, C.c_size_t, szp)
    f("trim_many", vp, C.c_size_t, vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_size_t, vp)
    f("stopping", vp, C.c_int, C.c_size_t, vp, vp, vp, vp)
    f("destroy", vp)
    f("get_tables", vp, vp, vp, vp)
    del u64p


def default_config(**kw):
    cfg = Config()
    cfg.tmin, cfg.tau, cfg.cw, cfg.length_scale = 0.2, 0.0, 0.001, 1.0
    cfg.potential, cfg.follow, cfg.follow_max_gen, cfg.vacancy_model = POT_UNIVERSAL, FOLLOW_ALL, 1, VAC_COUNT
    cfg.vmap_z[0] = cfg.vmap_z[1] = cfg.vmap_z[2] = -1
    for k, v in kw.items():
        if k == "vmap_z":
            for i in range(3):
                cfg.vmap_z[i] = v[i]
        else:
            if not hasattr(cfg, k):
                raise AttributeError(k)
            setattr(cfg, k, v)
    return cfg


def pack_materials(materials):
    """materials: list of dicts {rho, tag?, elements: [{Z, m, t, Edisp?, Elbind?}, ...]}"""
    mats = (Material * len(materials))()
    nel = sum(len(m["elements"]) for m in materials)
    els = (Element * nel)()
    k = 0
    for i, m in enumerate(materials):
        mats[i].rho = m["rho"]
        mats[i].tag = m.get("tag", -1)
        mats[i].n_elements = len(m["elements"])
        mats[i].first_element = k
        for e in m["elements"]:
            els[k].Z = e["Z"]
            els[k].m = e["m"]
            els[k].t = e["t"]
            els[k].Edisp = e.get("Edisp", 25.0)   # element.C:25
            els[k].Elbind = e.get("Elbind", 3.0)
            k += 1
    return mats, els


def pack_geometry(kind, w, bc=(BC_PBC, BC_PBC, BC_PBC), layers=None, kn=None, clusters=None):
    g = Geometry()
    g.kind = kind
    keep = []
    for i in range(3):
        g.w[i] = w[i]
        g.bc[i] = bc[i]
    if layers is not None:
        arr = np.ascontiguousarray(layers, dtype=np.float64)
        g.n_layers = len(arr)
        g.layer_thickness = arr.ctypes.data_as(C.POINTER(C.c_double))
        keep.append(arr)
    if kn is not None:
        for i in range(3):
            g.kn[i] = kn[i]
    if clusters is not None:
        arr = np.ascontiguousarray(clusters, dtype=np.float64).reshape(-1, 4)
        g.n_clusters = len(arr)
        g.cluster_xyzr = arr.ctypes.data_as(C.POINTER(C.c_double))
        keep.append(arr)
    return g, keep


def make_ions(n, Z, m, E, pos=(0.0, 50.0, 50.0), direction=(1.0, 0.0, 0.0), Ef=3.0, gen=0, tag=-1, seeds=None):
    """n identical primaries (runmytrim.C:291-304)."""
    ions = np.zeros(n, dtype=ION_DTYPE)
    ions["pos"] = pos
    ions["dir"] = direction
    ions["E"] = E
    ions["m"] = m
    ions["Ef"] = Ef
    ions["Z"] = Z
    ions["gen"] = gen
    ions["tag"] = tag
    if seeds is not None:
        ions["seed"] = seeds
    return ions


def fission_pairs(seed, first_event, n_events, w):
    """mtb_fission_pairs: the fission-fragment source of the UO2 experiment (2 primaries per event)."""
    lib = load_library()
    lib.mtb_fission_pairs.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.c_void_p,
                                      C.POINTER(C.c_double)]
    lib.mtb_fission_pairs.restype = C.c_int
    ions = np.zeros(2 * n_events, dtype=ION_DTYPE)
    box = (C.c_double * 3)(*w)
    e = C.c_double()
    rc = lib.mtb_fission_pairs(seed, first_event, n_events, box, ions.ctypes.data, C.byref(e))
    if rc != OK:
        raise MytrimError(rc, "mtb_fission_pairs")
    return ions


class EngineBase:
    """Call sequence of the reference apps over a C library with the mtb_ entry-point shapes."""

    _prefix = None

    def __init__(self, lib, handle):
        self._lib = lib
        self._h = handle
        self._keep = []

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)

    def _check(self, rc):
        if rc != OK:
            raise MytrimError(rc, self._error_text())
        return rc

    def _error_text(self):
        return ""

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_materials(self, materials):
        mats, els = pack_materials(materials)
        self._check(self._fn("set_materials")(self._h, len(mats), mats, len(els), els))

    def set_geometry(self, kind, w, **kw):
        g, keep = pack_geometry(kind, w, **kw)
        self._check(self._fn("set_geometry")(self._h, C.byref(g)))

    def set_layers(self, thicknesses, wy=100.0, wz=100.0, wx=None):
        """SampleLayers(thickness, wy, wz) as built by runmytrim (runmytrim.C:184)."""
        t = list(thicknesses)
        self.set_geometry(GEOM_LAYERS, (wx if wx is not None else float(sum(t)), wy, wz), layers=t)

    def run(self, ions, seed, first_index=0, records=False, check=True):
        ions = np.ascontiguousarray(ions, dtype=ION_DTYPE)
        rec = np.zeros(len(ions), dtype=RECORD_DTYPE) if records else None
        rc = self._fn("run")(self._h, len(ions), ions.ctypes.data, seed, first_index,
                             rec.ctypes.data if records else None)
        if check:
            self._check(rc)
        return rec

    def reset_tallies(self):
        self._check(self._fn("reset_tallies")(self._h))

    def counters(self):
        c = Counters()
        self._check(self._fn("get_counters")(self._h, C.byref(c)))
        return c.as_dict()

    def vac_depth(self, capacity=1 << 20):
        vac = np.zeros(capacity, dtype=np.uint64)
        repl = np.zeros(capacity, dtype=np.uint64)
        n = C.c_size_t()
        self._check(self._fn("get_vac_depth")(self._h, vac.ctypes.data, repl.ctypes.data, capacity, C.byref(n)))
        return vac[:n.value].copy(), repl[:n.value].copy()

    def vac_energy(self, rows=32, bins=1024):
        out = np.zeros((rows, bins), dtype=np.uint64)
        self._check(self._fn("get_vac_energy")(self._h, out.ctypes.data, rows, bins))
        return out

    def vacmap(self):
        out = np.zeros((VMAP_NX, VMAP_NY, 3), dtype=np.uint64)
        self._check(self._fn("get_vacmap")(self._h, out.ctypes.data))
        return out

    def ion_log(self, capacity=1 << 20):
        out = np.zeros(capacity, dtype=IONLOG_DTYPE)
        n = C.c_size_t()
        self._check(self._fn("get_ion_log")(self._h, out.ctypes.data, capacity, C.byref(n)))
        return out[:n.value].copy()

    def trim_one(self, ion, seed, uid, capacity=1 << 16):
        ion = np.array(ion, dtype=ION_DTYPE).reshape(1).copy()
        ev = np.zeros(capacity, dtype=EVENT_DTYPE)
        n = C.c_size_t()
        st = C.c_int32()
        self._check(self._fn("trim_one")(self._h, ion.ctypes.data, seed, uid, C.byref(st), ev.ctypes.data,
                                         capacity, C.byref(n)))
        return ion[0], st.value, ev[:n.value].copy()


def _trim_many(self, ions, seed, first_uid, events_per_ion, uids=None):
    """mtb_trim_many: every ion of the batch in one launch; returns (final ions, final states, counts, events[n][K])."""
    ions = np.array(ions, dtype=ION_DTYPE).copy()
    n = len(ions)
    ev = np.zeros((n, events_per_ion), dtype=EVENT_DTYPE)
    counts = np.zeros(n, dtype=np.uint32)
    states = np.zeros(n, dtype=np.int32)
    u = None if uids is None else np.ascontiguousarray(uids, dtype=np.uint64)
    self._check(self._fn("trim_many")(self._h, n, ions.ctypes.data, seed, first_uid, None if u is None else u.ctypes.data,
                                      states.ctypes.data, ev.ctypes.data, events_per_ion, counts.ctypes.data))
    return ions, states, counts, ev


EngineBase.trim_many = _trim_many


class Engine(EngineBase):
    """The CUDA engine (libmytrim_b200.so).  Raises if the library or a B200-class GPU is absent."""

    _prefix = "mtb_"

    def __init__(self, config=None, **kw):
        lib = load_library()
        cfg = config if config is not None else default_config(**kw)
        h = C.c_void_p()
        lib.mtb_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        lib.mtb_create.restype = C.c_int
        rc = lib.mtb_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise MytrimError(rc, lib.mtb_last_error().decode())
        super().__init__(lib, h)
        self.config = cfg

    def _error_text(self):
        return self._lib.mtb_last_error().decode()

    def run_beam(self, n, ion, seed, first_index=0, records=False):
        ion = np.array(ion, dtype=ION_DTYPE).reshape(1)
        rec = np.zeros(n, dtype=RECORD_DTYPE) if records else None
        self._check(self._lib.mtb_run_beam(self._h, n, ion.ctypes.data, seed, first_index,
                                           rec.ctypes.data if records else None))
        return rec

    def upload_primaries(self, ions):
        ions = np.ascontiguousarray(ions, dtype=ION_DTYPE)
        self._check(self._lib.mtb_upload_primaries(self._h, len(ions), ions.ctypes.data))

    def upload_primaries_ptr(self, n, host_ptr):
        self._check(self._lib.mtb_upload_primaries(self._h, n, host_ptr))

    def launch_resident(self, seed, first_index=0):
        self._check(self._lib.mtb_launch_resident(self._h, seed, first_index))

    def synchronize(self):
        self._check(self._lib.mtb_synchronize(self._h))

    def kernel_variant(self):
        """Name of the compile-time kernel variant this configuration selects (mtb_kernel_variant)."""
        fn = self._lib.mtb_kernel_variant
        fn.argtypes = [C.c_void_p]
        fn.restype = C.c_char_p
        return fn(self._h).decode()

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self._lib.mtb_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def fetch_records(self, n):
        rec = np.zeros(n, dtype=RECORD_DTYPE)
        self._check(self._lib.mtb_fetch_records(self._h, n, rec.ctypes.data))
        return rec

    def range_list(self, capacity=1 << 22):
        x = np.zeros(capacity, dtype=np.float32)
        z = np.zeros(capacity, dtype=np.int32)
        n = C.c_size_t()
        self._check(self._lib.mtb_get_range_list(self._h, x.ctypes.data, z.ctypes.data, capacity, C.byref(n)))
        return x[:n.value].copy(), z[:n.value].copy()

    def stopping(self, material, Z1, m1, E):
        Z1 = np.ascontiguousarray(Z1, dtype=np.int32)
        m1 = np.ascontiguousarray(m1, dtype=np.float64)
        E = np.ascontiguousarray(E, dtype=np.float64)
        out = np.zeros(len(E), dtype=np.float64)
        self._check(self._lib.mtb_stopping(self._h, material, len(E), Z1.ctypes.data, m1.ctypes.data,
                                           E.ctypes.data, out.ctypes.data))
        return out

    def tally_device_views(self):
        pu, pf = C.c_void_p(), C.c_void_p()
        nu, nf = C.c_size_t(), C.c_size_t()
        self._check(self._lib.mtb_tally_device_views(self._h, C.byref(pu), C.byref(nu), C.byref(pf), C.byref(nf)))
        return pu.value, nu.value, pf.value, nf.value
