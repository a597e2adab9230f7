#!/bin/bash
# compute-sanitizer passes over a small run of every kernel variant (race detection on
# the tally/queue code).  Usage (under gpurun): bash tools/sanitize.sh [tag]
TAG=${1:-r01}
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
PY
for tool in memcheck racecheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/sanitize_run.py > $OUT/${TAG}_sanitizer_$tool.log 2>&1
  tail -4 $OUT/${TAG}_sanitizer_$tool.log
done
