#!/bin/bash
# same-box A/B/C: current build, current without the cached dropout bits, c37 build
mkdir -p gpurun_out
O=gpurun_out/r02_c44
run() { n=$1; shift; env "$@" timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_$n.json 2> ${O}_$n.err; }
run cur LFMQ_X=0
run nodmask LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_alt.so
run c37 LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_c37.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_cur.csv python tools/run_once.py --workload cfg3 --steps 1 > /dev/null 2>&1
LFMQ_LIB_PATH=$PWD/lfm_quant_b200/_lfmq_c37.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_c37.csv python tools/run_once.py --workload cfg3 --steps 1 > /dev/null 2>&1
python - <<'PY'
import json,csv,collections
for n in ('cur','nodmask','c37'):
    try:
        d=json.loads(open('gpurun_out/r02_c44_%s.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['roofline']['regions_ms_per_step'].items()})
    except Exception as e:
        print(n, 'ERR', e)
def load(f):
    rows=list(csv.reader(open(f)))
    hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr=rows[hi]; col={n:i for i,n in enumerate(hdr)}
    agg=collections.OrderedDict()
    for r in rows[hi+1:]:
        if len(r)<len(hdr): continue
        k=r[col['Kernel Name']].split('(')[0][:50]
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[col['Metric Value']].replace(',',''))/1000
    return agg
a=load('gpurun_out/r02_c44_launches_cur.csv'); b=load('gpurun_out/r02_c44_launches_c37.csv')
for k in a:
    print('%-52s cur %3d %9.1f us   c37 %3d %9.1f us'%(k,a[k][0],a[k][1],b.get(k,[0,0])[0],b.get(k,[0,0])[1]))
PY
