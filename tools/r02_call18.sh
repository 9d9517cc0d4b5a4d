#!/bin/bash
# direction / x-update kernels: unroll and occupancy variants (build/variants/*.so), ncu durations of 24 launches each
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --no-cpu --no-profile"
for so in default build/variants/*.so; do
  name=$(basename $so .so)
  if [ "$so" = default ]; then unset DFDM_B200_LIB; else export DFDM_B200_LIB=$PWD/$so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pcg_direction|k_pcg_xupdate" -s 40 -c 24 --csv --log-file gpurun_out/r02_dirx_$name.csv $CMD > gpurun_out/ncu_d.log 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r02_dirx_$name.csv")) if len(r)>10]
hdr=rows[0]; i_k=hdr.index("Kernel Name"); i_v=hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]: d[r[i_k].split("(")[0]].append(float(r[i_v].replace(",",""))/1e3)
print("$name", {k:[round(x,1) for x in sorted(v)[-3:]] + [len(v)] for k,v in d.items()})
PY
done
