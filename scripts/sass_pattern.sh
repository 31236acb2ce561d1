#!/bin/bash
# instruction-class map of one kernel's SASS around its DMMA stream (M = DMMA, f = FP64 pipe, l = LDS, | = control flow)
# usage: scripts/sass_pattern.sh <object file> <substring of the mangled kernel name>
obj=$1; pat=$2
cuobjdump -sass $obj | awk -v pat="$pat" '/Function :/{on = index($0, pat) > 0} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" > /tmp/_k.sass
first=$(grep -n DMMA /tmp/_k.sass | head -1 | cut -d: -f1); last=$(grep -n DMMA /tmp/_k.sass | tail -1 | cut -d: -f1)
echo "instructions $(wc -l < /tmp/_k.sass), DMMA $(grep -c DMMA /tmp/_k.sass), map of lines $((first-40))..$((last+40))"
sed -n "$((first-40)),$((last+40))p" /tmp/_k.sass | awk '{ins=$2; sub(/;.*/,"",ins); if(ins ~ /^@/){ins=$3}; split(ins,a,"."); k=a[1]; c=(k=="DMMA")?"M":(k ~ /^D(FMA|ADD|MUL|SETP)/)?"f":(k=="LDS")?"l":(k ~ /^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|CALL|RET)/)?"|":"."; printf "%s", c} END{print ""}' | fold -w 120
