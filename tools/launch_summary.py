import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr = rows[hi]; data = rows[hi+1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict(); tot=0
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*','',r[ki])[:72]
    v = float(r[vi].replace(',',''))
    if r[ui]=='us': v*=1e3
    elif r[ui]=='ms': v*=1e6
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=v; tot+=v
print('total ms %.3f launches %d' % (tot/1e6, len(data)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:n]:
    print(f'{t/1e6:9.3f} ms {100*t/tot:5.1f}% x{c:4d}  {k}')
