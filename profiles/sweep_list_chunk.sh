# (historic: TMD_NL_CHUNK was a knob of the r01o list kernel; chunks are fixed at 16 since r01q)
for c in 8 16 24; do
  TMD_NL_CHUNK=$c timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nl_list_kernel -c 12 --csv --log-file gpurun_out/list_chunk$c.csv python bench.py --quick --steps 3 --warmup 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/list_chunk$c.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; start=i+1; break
vi=hdr.index("Metric Value")
v=[float(r[vi].replace(",","")) for r in rows[start:] if len(r)>vi]
print("chunk $c: real builds (us):", [round(x/1e3,1) for x in v if x>50000])
PY
done
