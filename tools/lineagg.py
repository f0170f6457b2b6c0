# aggregate ncu per-SASS metrics by source line using nvdisasm -g line info (instruction order match)
import csv, re, sys, collections
sass_file, csv_file, kernel_tag, src_file = sys.argv[1:5]
lines=open(sass_file).read().split('\n')
# find the section of the kernel
start=[i for i,l in enumerate(lines) if l.startswith('.text.') and kernel_tag in l or (l.strip().startswith('.section') and '.text.' in l and kernel_tag in l)]
s=start[0]
cur=None; seq=[]
for l in lines[s+1:]:
    if l.strip().startswith('.section') and '.text.' in l: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)))
        # inlined-at info: keep the innermost
        continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: seq.append((cur,m.group(2)))
rows=list(csv.reader(open(csv_file)))
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address']
k=int(sys.argv[5]) if len(sys.argv)>5 else 0
h=rows[hi[k]]; end=hi[k+1]-1 if k+1<len(hi) else len(rows)
body=rows[hi[k]+1:end]
print('sass instrs',len(seq),'csv instrs',len(body))
ci=h.index('# Samples'); ii=h.index('Instructions Executed'); ti=h.index('Thread Instructions Executed')
agg=collections.defaultdict(lambda:[0,0,0])
for (loc,txt),r in zip(seq,body):
    a=agg[loc]; a[0]+=int(r[ci]); a[1]+=int(r[ii]); a[2]+=int(r[ti])
src={}
tots=[sum(a[i] for a in agg.values()) for i in range(3)]
print('totals samples/inst/threadinst',tots)
for loc,a in sorted(agg.items(), key=lambda kv:(kv[0][0] if kv[0] else '', kv[0][1] if kv[0] else 0)):
    print(f"{str(loc):34s} samples {a[0]:6d} ({100*a[0]/tots[0]:5.1f}%)  winst {a[1]:10d} ({100*a[1]/tots[1]:5.1f}%)")
