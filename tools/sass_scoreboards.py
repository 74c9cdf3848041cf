"""SASS control words of a kernel in libidcodec.so: write / read barrier and wait mask per instruction; prints every LDG and every
instruction that waits for a scoreboard an LDG uses. usage: python tools/sass_scoreboards.py <substring of the mangled kernel name>"""
import re, sys, subprocess
so='/root/repo/vector_db_id_compression_b200/libidcodec.so'
pat=sys.argv[1]
txt=subprocess.run(["cuobjdump","-sass",so],capture_output=True,text=True).stdout
# split functions
funcs=re.split(r'\n\s*Function : ', txt)
for f in funcs:
    name=f.split('\n',1)[0]
    if pat not in name: continue
    lines=f.split('\n'); ins=[]; i=0
    while i<len(lines):
        m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/',lines[i])
        if m and i+1<len(lines):
            m2=re.match(r'\s+/\* (0x[0-9a-f]{16}) \*/',lines[i+1])
            if m2:
                hi=int(m2.group(1),16); ctrl=(hi>>41)&0x7fffff
                ins.append((int(m.group(1),16),m.group(2),(ctrl>>5)&7,(ctrl>>8)&7,(ctrl>>11)&0x3f)); i+=2; continue
        i+=1
    print("==",name[:90], len(ins),"instructions")
    key=sys.argv[2] if len(sys.argv)>2 else 'LDG'
    for ad,t,wbar,rbar,wait in ins:
        if key in t or wait:
            pass
    # print every LDG-class instruction and every waiter on any barrier an LDG uses
    ldg_bars=set(w for ad,t,w,r,wa in ins if ('LDG' in t and 'LDGSTS' not in t and 'DEPBAR' not in t) and w!=7)
    print("barriers used by LDGs:",ldg_bars)
    for ad,t,wbar,rbar,wait in ins:
        if ('LDG' in t and 'LDGSTS' not in t) or any((wait>>b)&1 for b in ldg_bars):
            print(hex(ad), t[:80].ljust(80), "W%s"%(wbar if wbar!=7 else '-'),"R%s"%(rbar if rbar!=7 else '-'),"wait",format(wait,'06b'))
