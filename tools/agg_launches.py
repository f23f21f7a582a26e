import csv,collections,sys
def agg(path, top=16):
    with open(path) as f:
        lines=[l for l in f if not l.startswith('==')]
    r=csv.DictReader(lines)
    a=collections.defaultdict(lambda:[0,0.0]); tot=0
    for row in r:
        if row.get('Metric Name')!='gpu__time_duration.sum': continue
        name=row['Kernel Name'].split('(')[0].replace('void ','')
        v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v*={'ns':1,'us':1e3,'ms':1e6,'s':1e9}.get(u,1)
        a[name][0]+=1; a[name][1]+=v; tot+=v
    print(path,'total ms %.3f'%(tot/1e6))
    for k,(n,t) in sorted(a.items(), key=lambda x:-x[1][1])[:top]:
        print('  %-64s n=%4d %8.3f ms %5.1f%% avg %7.1f us'%(k[:64],n,t/1e6,100*t/tot,t/n/1e3))
for p in sys.argv[1:]: agg(p)
