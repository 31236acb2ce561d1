#!/bin/bash
# ncu --set full capture of the dominant kernel of every BASELINE config (small chain counts); run under gpurun.
# Only the raw-metric CSV pages travel back (the .ncu-rep files with imported source exceed the 64 MiB limit).
mkdir -p gpurun_out
N="ncu --set full --clock-control none -c 1"
run() { # name kernel-regex command...
  name=$1; rx=$2; shift 2
  $N -k regex:$rx -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
run prof_c3_funnel walnutspy python scripts/config_sweep.py --only c3 --scale 0.25
run prof_c4_logreg walnutspy python scripts/logreg_bench.py --reps 1 --integrator R2P --chains 1184
run prof_c5_sw walnutspy python scripts/config_sweep.py --only c5 --scale 0.03
run prof_c1_package package python scripts/config_sweep.py --only c1 --scale 1
ls -la gpurun_out
