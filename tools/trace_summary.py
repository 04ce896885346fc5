import re, sys
ev=[]
for l in open(sys.argv[1]):
    m=re.match(r'\s*(\d+) (\w+)\s+(\w) tile(\d) stage\s*(\d+) item\s*(\d+)',l)
    if m: ev.append((int(m[1]),m[2],m[3],int(m[5]),int(m[6])))
st={}
for clk,role,kind,stage,item in ev:
    d=st.setdefault(stage,{})
    if role=='mma' and kind=='A': d['A']=clk
    if role=='mma' and kind=='F': d.setdefault('F',[]).append(clk)
    if role=='mma' and kind=='C': d['C']=clk
    if role.startswith('epi') and kind=='W': d['W']=clk
    if role.startswith('epi') and kind=='D': d['D']=clk
prevD=None; first=None
for s in sorted(st):
    d=st[s]; F=d.get('F',[])
    if first is None: first=d['A']
    print("stage %2d: aready@%7d loads %2d firstF +%5d lastF +%5d commit +%5d | acc seen +%5d epi done +%5d (epi %5d) handoff %s"%(
        s,d['A']-first,len(F),F[0]-d['A'],F[-1]-d['A'],d['C']-d['A'],d['W']-d['A'],d['D']-d['A'],d['D']-d['W'],(d['A']-prevD) if prevD else None))
    prevD=d['D']
print("tile total", prevD-first)
