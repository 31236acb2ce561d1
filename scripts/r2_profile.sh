#!/bin/bash
# ncu --set full captures (with CUDA source correlation) of the shipped kernels; run under gpurun.
#   scripts/r2_profile.sh <tag> <name>...      names: c2 c2nuts c3 c3nuts c3lone c3neck c5 c4 c1
# Per capture: gpurun_out/<tag>_<name>.raw.csv (raw metrics) and .src.csv (per-line stall samples); the .ncu-rep
# files stay on the box (too large to travel).
tag=$1; shift
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -s 1 -c 1"
cap() { # name kernel-regex command...
  name=$1; rx=$2; shift 2
  $N -k regex:$rx -o /tmp/${tag}_$name "$@" > gpurun_out/${tag}_$name.log 2>&1
  ncu -i /tmp/${tag}_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_$name.raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_$name.src.csv 2>/dev/null
  rm -f /tmp/${tag}_$name.ncu-rep
}
for what in "$@"; do
  case $what in
    c2)     cap c2 walnutspy python scripts/config_sweep.py --only c2 --integ R2P --scale 0.0723 ;;      # 4738 chains
    c2nuts) cap c2nuts walnutspy python scripts/config_sweep.py --only c2 --integ fixed --scale 0.0723 ;;
    c3)     cap c3 walnutspy python scripts/config_sweep.py --only c3 --integ R2P --scale 0.125 ;;
    c3nuts) cap c3nuts walnutspy python scripts/config_sweep.py --only c3 --integ fixed --scale 0.125 ;;
    c5)     cap c5 walnutspy python scripts/config_sweep.py --only c5 --integ R2P --scale 0.03 ;;
    c4)     cap c4 walnutspy python scripts/config_sweep.py --only c4 --integ R2P --scale 0.0723 ;;
    c3lone) N="ncu --set full --clock-control none --import-source on -c 1"
            cap c3lone walnutspy python scripts/funnel_tail.py 65536 10125 ;;   # the costliest chain of C3 (funnel mouth) ALONE
    c3neck) N="ncu --set full --clock-control none --import-source on -c 1"
            cap c3neck walnutspy python scripts/funnel_tail.py 262144 222477 ;; # the costliest neck chain ALONE
    c1)     cap c1 package python scripts/config_sweep.py --only c1 --scale 1 ;;
  esac
done
ls -la gpurun_out | tail -20
