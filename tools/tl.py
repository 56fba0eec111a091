import re, sys
lines=open(sys.argv[1]).read().splitlines()[2:]
k=int(sys.argv[2]) if len(sys.argv)>2 else 3
ev=[]
for l in lines:
    m=re.match(r'\s*(\d+) \(\+\s*\d+\)\s+(.*)',l)
    if m: ev.append((int(m.group(1)), m.group(2).strip()))
ts0=[c for c,s in ev if s.startswith('slot0 epi') and 'tile start' in s]
ts1=[c for c,s in ev if s.startswith('slot1 epi') and 'tile start' in s]
print(ts0, [b-a for a,b in zip(ts0,ts0[1:])]); print(ts1, [b-a for a,b in zip(ts1,ts1[1:])])
a,b=ts0[k],ts0[k+1]
prev={}
for c,s in ev:
    if a<=c<=b and not s.startswith('MMA'):
        sl=s[:5]
        print(f"{c-a:7d} (+{c-prev.get(sl,a):5d}) {'' if sl=='slot0' else '                              '}{s}"); prev[sl]=c
