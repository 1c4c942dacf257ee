# device time of the real list builds (ncu, serialised): python bench.py --quick --steps 3 --warmup 3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nl_list_kernel|nl_bin_kernel" -c 24 --csv --log-file gpurun_out/list_build.csv python bench.py --quick --steps 3 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/list_build.csv")))
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; start=i+1; break
vi=hdr.index("Metric Value"); ki=hdr.index("Kernel Name")
for name in ("nl_list_kernel","nl_bin_kernel"):
    v=[float(r[vi].replace(",","")) for r in rows[start:] if len(r)>vi and name in r[ki]]
    print(name, "real builds (us):", [round(x/1e3,1) for x in v if x>30000], "gated (us):", [round(x/1e3,1) for x in v if x<=30000][:4])
PY
