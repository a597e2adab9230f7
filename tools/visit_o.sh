# Round-2 visit o: resident CTAs per SM of the FAST + share kernel (Xe->UO2 10 MeV, Xe->ZrO2)
OUT=gpurun_out; LOG=$OUT/r02o2_fastshare.log; : > $LOG
for v in cur13 fs5 fs7 cur13; do
  for w in "xe_on_uo2_10MeV 8192 1" "xe_on_zro2_500keV 65536 1" "xe_on_zro2_500keV 131072 1"; do
    set -- $w
    echo "== $v $1 n=$2" >> $LOG
    MYTRIM_B200_LIB=$PWD/build/variants/$v.so timeout 200 python tools/profile_run.py --workload $1 --primaries $2 --tally $3 --launches 3 2>&1 | tail -1 >> $LOG
  done
done
cat $LOG
