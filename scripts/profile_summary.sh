#!/bin/bash
# profiles/<out>.txt from one capture of scripts/r2_profile.sh: raw-metric summary + the heaviest CUDA source lines
# usage: scripts/profile_summary.sh gpurun_out/<tag>_<name> profiles/<out>.txt "<header line>"
base=$1; out=$2; shift 2
{
  echo "# $*"
  python scripts/summarize_ncu.py $base.raw.csv - /tmp/_sum.txt > /dev/null; grep -v "^$" /tmp/_sum.txt
  echo
  echo "# per CUDA source line (ncu --page source --print-source cuda,sass; scripts/ncu_lines.py): share of warp-stall samples, share of executed warp instructions, dominant stall reasons"
  python scripts/ncu_lines.py $base.src.csv 28 | cut -c1-190
} > $out
