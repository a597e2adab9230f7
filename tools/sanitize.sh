#!/bin/bash
# compute-sanitizer passes over a small run of every kernel variant (race detection on
# the tally/queue code).  Usage (under gpurun): bash tools/sanitize.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/sanitize_run.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from mytrim_b200 import capi
from tests import util
# fast + share kernel (small launch), generic + share (phonon), events, stopping, clusters/custom species
for mask in (capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS, capi.TALLY_PHONON | capi.TALLY_RECORDS | capi.TALLY_IONLOG):
    with capi.Engine(tally_mask=mask) as eng:
        c = util.setup_engine(eng, "cu_on_cu_10keV")
        eng.run(util.primaries_for(c, 600), seed=1, records=True)
        print(mask, eng.counters()["steps"])
with capi.Engine(tally_mask=capi.TALLY_RECORDS) as eng:
    c = util.setup_engine(eng, "xe_on_zro2_500keV")
    eng.run(util.primaries_for(c, 4), seed=1, records=True)
    print("zro2", eng.counters()["steps"], eng.counters()["stack_max"])
    f, s, ev = eng.trim_one(util.primaries_for(c, 1)[0], 3, 7)
    print("events", len(ev))
    print(eng.stopping(0, [54], [131.0], [5e5]))
# layers variant (+share): follow policy, Kinchin-Pease estimate, 2-D tally
with capi.Engine(tally_mask=capi.TALLY_VAC_ENERGY | capi.TALLY_VAC_DEPTH | capi.TALLY_RECORDS, follow=capi.FOLLOW_GEN_LT,
                 follow_max_gen=2, vacancy_model=capi.VAC_KP) as eng:
    c = util.setup_engine(eng, "xe_on_zro2_500keV")
    eng.run(util.primaries_for(c, 8), seed=1, records=True)
    print("layers variant", eng.counters()["steps"])
# clusters variant (+share): bubbles, per-primary species, ion log; then CUT boundaries -> all-options kernel
cl = np.loadtxt("tests/golden/uo2_out.clcoor")[:, :4]
rng = np.random.default_rng(2)
ions = capi.make_ions(24, 1, 1.0, 1.0)
ions["Z"] = rng.integers(30, 62, 24)
ions["m"] = np.round(ions["Z"] * 2.55 + rng.uniform(-3, 3, 24), 3)
ions["E"] = rng.uniform(2e4, 2e5, 24)
ions["pos"] = rng.uniform(0, 400, (24, 3))
ions["pos"][:6] = cl[np.arange(6) % len(cl), :3] + 1.0
d = rng.normal(size=(24, 3))
ions["dir"] = d / np.linalg.norm(d, axis=1)[:, None]
ions_cl = ions.copy()
for bc in ((capi.BC_PBC,) * 3, (capi.BC_CUT, capi.BC_PBC, capi.BC_INF)):
    with capi.Engine(tally_mask=capi.TALLY_PHONON | capi.TALLY_RECORDS | capi.TALLY_IONLOG, ionlog_z=54) as eng:
        eng.set_materials([util.UO2, util.XE_GAS])
        eng.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), bc=bc, kn=(39, 39, 39), clusters=cl)
        eng.run(ions, seed=1, records=True)
        print("clusters", bc, eng.counters()["steps"], len(eng.ion_log()))
# plain (non-sharing) kernels: more primaries than 4 per lane is too slow under the sanitizer, so sharing is switched off
import os
os.environ["MYTRIM_B200_NO_SHARE"] = "1"
with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH) as eng:
    c = util.setup_engine(eng, "cu_on_cu_1keV")
    eng.run(util.primaries_for(c, 4096), seed=1)                      # MONO no-records (what bench.py times)
    print("plain MONO, no records", eng.counters()["steps"])
    eng.run(util.primaries_for(c, 2048), seed=2, records=True)        # MONO with records
    print("plain MONO, records", eng.counters()["steps"])
with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH) as eng:
    c = util.setup_engine(eng, "xe_on_zro2_500keV")                   # compound stack folded into one material: FAST
    ions = util.primaries_for(c, 16)
    ions["E"] = 2.0e4
    eng.run(ions, seed=1)
    print("plain FAST", eng.counters()["steps"])
del os.environ["MYTRIM_B200_NO_SHARE"]
with capi.Engine(tally_mask=capi.TALLY_VAC_DEPTH) as eng:
    c = util.setup_engine(eng, "cu_on_cu_10keV")
    eng.run(util.primaries_for(c, 512), seed=1)                       # MONO + share
    print("MONO + share", eng.counters()["steps"])
# compile-time tally variants: MONO-EVAC (+share), FAST-PHONON (+share), CLUSTERS-LOG (+share)
with capi.Engine(tally_mask=capi.TALLY_VAC_ENERGY) as eng:
    c = util.setup_engine(eng, "c_on_w_1MeV")
    ions = util.primaries_for(c, 64)
    ions["E"] = 5.0e4
    eng.run(ions, seed=3)
    print("MONO-EVAC + share", eng.counters()["steps"], int(eng.vac_energy().sum()))
with capi.Engine(tally_mask=capi.TALLY_PHONON) as eng:
    c = util.setup_engine(eng, "xe_on_zro2_500keV")
    ions = util.primaries_for(c, 16)
    ions["E"] = 2.0e4
    eng.run(ions, seed=3)
    print("FAST-PHONON + share", eng.counters()["steps"], eng.counters()["EnucTotal"])
with capi.Engine(tally_mask=capi.TALLY_IONLOG, ionlog_z=54) as eng:
    eng.set_materials([util.UO2, util.XE_GAS])
    eng.set_geometry(capi.GEOM_CLUSTERS, (400.0, 400.0, 400.0), kn=(39, 39, 39), clusters=cl)
    eng.run(ions_cl, seed=3)
    print("CLUSTERS-LOG + share", eng.counters()["steps"], len(eng.ion_log()))
with capi.Engine(tally_mask=0) as eng:
    c = util.setup_engine(eng, "cu_on_cu_10keV")
    fin, st, cnt, ev = eng.trim_many(util.primaries_for(c, 100), 5, 0, 64)   # event mode, one lane per ion
    print("trim_many", int(cnt.sum()))
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_run.py > $OUT/${TAG}_sanitizer_$tool.log 2>&1
  tail -4 $OUT/${TAG}_sanitizer_$tool.log
done
