mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 60 --csv --log-file gpurun_out/r1_j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/r1_j_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; k=h.index('Kernel Name'); v=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hdr+1:]:
    if len(r)>v:
        n=r[k].split('(')[0]; agg[n][0]+=1; agg[n][1]+=float(r[v].replace(',',''))
tot=sum(x[1] for x in agg.values())
for n,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:10]: print(f"{n[:60]:60s} launches {c:4d} total {t/1e3:10.1f} us share {t/tot:.3f}")
PY
