#!/bin/bash
# launch list of one cfg3 train step (persistent steps): where the non-GEMM time goes
mkdir -p gpurun_out
O=gpurun_out/r02_c41
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file ${O}_launches_cfg3.csv python tools/run_once.py --workload cfg3 --steps 2 > ${O}_ncu.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_c41_launches_cfg3.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; col={n:i for i,n in enumerate(hdr)}
per=collections.OrderedDict()
ids={}
for r in rows[hi+1:]:
    if len(r)<len(hdr): continue
    k=(r[col['ID']], r[col['Kernel Name']].split('(')[0][:58])
    d=ids.setdefault(k,{})
    v=float(r[col['Metric Value']].replace(',',''))
    u=r[col['Metric Unit']]
    m=r[col['Metric Name']]
    if m=='gpu__time_duration.sum': d['us']=v/1000 if u.startswith('n') else v
    else:
        mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
        d['bytes']=d.get('bytes',0)+v*mult
n=len(ids)
keys=list(ids)
half=keys[len(keys)//2:]            # second step only
agg=collections.OrderedDict()
for k in half:
    a=agg.setdefault(k[1],[0,0.0,0.0]); a[0]+=1; a[1]+=ids[k].get('us',0); a[2]+=ids[k].get('bytes',0)
tot=sum(a[1] for a in agg.values())
print('second step: %d launches, %.0f us serialised'%(len(half),tot))
for name,a in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print('%-60s %4d %9.1f us %5.1f%%  %8.1f MB  %6.0f GB/s'%(name,a[0],a[1],100*a[1]/tot,a[2]/1e6,a[2]/1e3/max(a[1],1e-9)))
PY
