#!/bin/bash
# usage: tools/sass_seq.sh <object> <mangled-name-substring>: instruction-class sequence of one kernel
# (M = DMMA, f = scalar FP64, l = LDS, s = SHFL, B = branch, S = store/sync, . = other)
cuobjdump -sass "$1" 2>/dev/null | awk -v pat="$2" '/Function :/{f=($0 ~ pat)} f{print}' > /tmp/_k.sass
grep -cE "DMMA" /tmp/_k.sass | sed 's/^/DMMA: /'
grep -cE "DFMA|DMUL|DADD" /tmp/_k.sass | sed 's/^/scalar FP64: /'
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass | awk '{ if ($0 ~ /DMMA/) t="M"; else if ($0 ~ /DFMA|DMUL|DADD/) t="f"; else if ($0 ~ /LDS/) t="l"; else if ($0 ~ /BRA|EXIT|CALL|RET/) t="B"; else if ($0 ~ /SHFL/) t="s"; else if ($0 ~ /STS|SYNCS|ATOM|RED|STG|BAR/) t="S"; else t="."; printf "%s", t } END {print ""}' | fold -w 200
