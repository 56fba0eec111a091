"""Static LDL / STL instructions of one kernel per source line (deepest frame in the .cu).  usage: spill_lines.py disasm.txt kernel_substr file.cu"""
import re, collections, sys
dis, kern, cu = sys.argv[1:4]
frames=[]; infn=False; fresh=True
cnt=collections.Counter()
for ln in open(dis):
    if ln.startswith('.text.'):
        infn = kern in ln; continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh: frames=[]; fresh=False
        frames.append((m.group(1).split('/')[-1], int(m.group(2))))
        continue
    m=re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        fresh=True
        txt=m.group(2)
        if 'LDL' in txt or 'STL' in txt:
            deep=next((f for f in frames if f[0]==cu.split('/')[-1]), ('?',0))
            cnt[(deep[1], 'LDL' if 'LDL' in txt else 'STL')]+=1
src=open(cu).read().splitlines()
for (l,k),n in sorted(cnt.items()):
    print(l,k,n, src[l-1].strip()[:100] if l>0 else '')
