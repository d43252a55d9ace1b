python tools/bench_sw.py 65536 > gpurun_out/sw_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/sw_launches.csv python tools/bench_sw.py 65536 > gpurun_out/sw_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/sw_launches.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
hdr=rows[h]; ik,iv=hdr.index("Kernel Name"),hdr.index("Metric Value")
for r in rows[h+1:h+25]:
    print(r[ik][:60], r[iv])
PY
