#!/bin/bash
# per-kernel conv times with the patch kernels' stores / MMAs / TMA loads disabled (HP3D_CONV_DEBUG bit mask 1/2/4)
TAG=${1:-r01x}
OUT=gpurun_out; mkdir -p $OUT
for d in ${DBG:-0 1 2 4 7}; do
  HP3D_CONV_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'conv_|stem2' -c 44 --csv --log-file $OUT/${TAG}_convdbg_$d.csv python tools/bench_encoder.py > /dev/null 2>&1
done
python - <<'PY'
import csv,glob,re,collections
for f in sorted(glob.glob('gpurun_out/*_convdbg_*.csv')):
    rows=list(csv.reader(open(f)))
    h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.OrderedDict()
    for r in rows[h+1:]:
        if len(r)>vi:
            k=re.sub(r'\(CUtensor.*','',r[ki]); a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[vi].replace(',',''))
    print(f, {k[-40:]:(v[0],round(v[1]/1e3,1)) for k,v in agg.items()})
PY
