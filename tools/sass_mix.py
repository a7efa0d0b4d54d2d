#!/usr/bin/env python
"""Opcode mix of a kernel's hot loop from the shipped library: `python tools/sass_mix.py <lib.so> <substring of the
demangled kernel name> [min LDS.128 in the loop]`.  The hot loop is the smallest backward-branch region holding at
least that many LDS.128 (decoder: 4 LUT loads per word; encoder: 1 vector load).  ALU-ish = LOP3/SHF/IADD3/ISETP/SEL/
PRMT/LEA/VIADD/I2FP/PLOP3/min-max/MOV; FMA-ish = IMAD/FFMA/FMUL (B300_MICROARCH.md: both pipes issue one warp
instruction per two cycles)."""
import re,subprocess,sys,collections
so=sys.argv[1]; pat=sys.argv[2]; need=int(sys.argv[3]) if len(sys.argv)>3 else 4
out=subprocess.run(['cuobjdump','-sass',so],capture_output=True,text=True).stdout
funcs=out.split('Function : ')
for f in funcs[1:]:
    name=f.split('\n',1)[0]
    dem=subprocess.run(['c++filt',name.strip()],capture_output=True,text=True).stdout.strip()
    if pat not in dem: continue
    lines=[]
    for l in f.split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
        if m: lines.append((int(m.group(1),16),m.group(2)))
    # find backward branches
    best=None
    for idx,(addr,ins) in enumerate(lines):
        m=re.search(r'BRA\s+(?:\w+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)',ins) or re.search(r'BRA.*0x([0-9a-f]+)',ins)
        if m:
            tgt=int(m.group(1),16)
            if tgt<addr:
                body=[i for a,i in lines if tgt<=a<=addr]
                n128=sum('LDS.128' in i for i in body)
                if n128>=need and (best is None or len(body)<len(best[2])):
                    best=(tgt,addr,body)
    print(dem[:110])
    if not best: print('  no loop found'); continue
    body=best[2]
    ALU=('LOP3','SHF','IADD3','ISETP','SEL','PRMT','LEA','VIADD','I2FP','PLOP3','FSETP','IMNMX','VIMNMX','FMNMX','BFE','BFI','SGXT','MOV ','CS2R','FLO','POPC','F2I','I2F','FADD','VOTE')
    FMA=('IMAD','FFMA','FMUL','HFMA')
    c=collections.Counter()
    for i in body:
        op=i.split()[0] if not i.startswith('@') else i.split()[1]
        base=op.split('.')[0]
        c[base]+=1
    alu=sum(v for k,v in c.items() if any(k.startswith(a.strip()) for a in ('LOP3','SHF','IADD3','ISETP','SEL','PRMT','LEA','VIADD','I2FP','PLOP3','VIMNMX','FMNMX','MOV','CS2R','FADD')))
    fma=sum(v for k,v in c.items() if k in ('IMAD','FFMA','FMUL'))
    print('  loop %x..%x instrs %d  alu-ish %d  fma-ish %d'%(best[0],best[1],len(body),alu,fma))
    print('  ',sorted(c.items(),key=lambda kv:-kv[1]))
