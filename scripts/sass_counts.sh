#!/bin/bash
# SASS evidence (no GPU needed): instruction counts per kernel of the shipped library --
#   DFMA (FP64 FMA pipe), DMMA (FP64 tensor cores, mma.sync.m8n8k4.f64), UBLKCP (cp.async.bulk = TMA engine),
#   SYNCS (mbarrier), REDUX (warp-wide integer reduction), BAR (block barriers), LDL/STL (local memory = spills)
# usage: scripts/sass_counts.sh > profiles/rNN_sass_counts.txt
LIB=${1:-walnuts_b200/_lib/libwalnuts_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a); counts of static instructions per kernel"
echo "# kernel | DFMA DMMA UBLKCP SYNCS REDUX BAR LDL STL"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ {fn=$3}
/ DFMA/ {a[fn]++} / DMMA/ {b[fn]++} /UBLKCP/ {c[fn]++} /SYNCS/ {d[fn]++} /REDUX/ {e[fn]++} / BAR\./ {f[fn]++} / LDL/ {g[fn]++} / STL/ {h[fn]++} {seen[fn]=1}
END {for (k in seen) if (k != "") printf "%s | %d %d %d %d %d %d %d %d\n", k, a[k], b[k], c[k], d[k], e[k], f[k], g[k], h[k]}' | c++filt | sort
