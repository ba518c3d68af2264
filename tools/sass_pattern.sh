#!/bin/bash
# usage: sass_pattern.sh <object> <mangled-substring>   prints D=DMMA n=NOP f=FP64 l=LDS s=STL/LDL(spill) .=other
fn=$(cuobjdump -sass $1 2>/dev/null | grep "Function :" | grep "$2" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$fn" $1 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/\/\*[0-9a-f]*\*\///g; s/^\s*//' | awk '{op=$1; if(op ~ /^@/) op=$2; print op}' | sed 's/\..*//' | python3 -c "
import sys
ops=[l.strip() for l in sys.stdin]
m={'DMMA':'D','DFMA':'f','DMUL':'f','DADD':'f','NOP':'n','LDS':'l','STL':'s','LDL':'s','STS':'S','SYNCS':'y','BRA':'b','ATOMS':'a'}
s=''.join(m.get(o,'.') for o in ops)
print(len(ops),'instructions')
for k in range(0,len(s),150): print(s[k:k+150])
"
