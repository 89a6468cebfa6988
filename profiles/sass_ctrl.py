#!/usr/bin/env python
"""Decode the scheduling control fields of sm_100a SASS (stall count, scoreboard set / waited for) from
`cuobjdump -sass` output - how the "all affinity loads share one scoreboard" finding of DESIGN.md section 4 was made.

    cuobjdump -sass -fun <mangled kernel name> cosa_b200/libcosa_b200.so > k.sass
    python profiles/sass_ctrl.py k.sass 'LDG|SYNCS'      # instructions matching the regex, or that wait on a scoreboard

Upper 64-bit word of an instruction: bits 41-44 stall count, 45 yield, 46-48 scoreboard set on completion (7 = none),
49-51 read scoreboard, 52-57 mask of scoreboards waited for.
"""
import re,sys
# decode Volta+ control fields from cuobjdump -sass output (two hex words per instruction)
lines=open(sys.argv[1]).read().split('\n')
pat=re.compile(r'^\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/')
pat2=re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
out=[]
i=0
while i<len(lines):
    m=pat.match(lines[i])
    if m and i+1<len(lines):
        m2=pat2.match(lines[i+1])
        if m2:
            hi=int(m2.group(1),16)
            stall=(hi>>41)&0xf; yld=(hi>>45)&1; wbar=(hi>>46)&7; rbar=(hi>>49)&7; wait=(hi>>52)&0x3f
            out.append((m.group(1),m.group(2),stall,yld,wbar,rbar,wait))
            i+=2; continue
    i+=1
sel=sys.argv[2] if len(sys.argv)>2 else None
for a,ins,stall,yld,wbar,rbar,wait in out:
    if sel and not re.search(sel,ins) and wait==0: continue
    print("%s st=%2d w=%s r=%s wait=%s  %s"%(a,stall, wbar if wbar!=7 else '-', rbar if rbar!=7 else '-', ''.join(str(b) for b in range(6) if wait>>b&1) or '-', ins[:90]))
