#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-profile"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_pcg_direction|k_pcg_xupdate" -s 40 -c 24 --csv --log-file gpurun_out/r02_dirx.csv $CMD > gpurun_out/ncu_d.log 2>&1; echo "rc $?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r02_dirx.csv")) if len(r)>10]
hdr=rows[0]; i_k=hdr.index("Kernel Name"); i_m=hdr.index("Metric Name"); i_v=hdr.index("Metric Value"); i_id=hdr.index("ID")
d=collections.OrderedDict()
for r in rows[1:]: d.setdefault((r[i_id],r[i_k].split("(")[0]),{})[r[i_m]]=float(r[i_v].replace(",",""))
for k,v in d.items(): print(k[1], "%.1f us" % (v["gpu__time_duration.sum"]/1e3), "%.0f MB" % ((v["dram__bytes_read.sum"]+v["dram__bytes_write.sum"])/1e6), v["launch__registers_per_thread"], "%.0f%% warps" % v["sm__warps_active.avg.pct_of_peak_sustained_active"])
PY
