#!/bin/bash
# r02n: Zipf after the row-pointer / c prefetch, split rows anywhere in a tile (ring / slab), adaptive apply probe
mkdir -p gpurun_out
echo "== pytest csr + all"; timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -6 | tee gpurun_out/r02n_pytest.txt
echo "== zipf / cfg2"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1; timeout 300 python scripts/zipf_probe.py cfg2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' --csv --log-file gpurun_out/r02n_zipf_launches.csv python scripts/zipf_probe.py cfg2_zipf > gpurun_out/r02n_zipf_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02n_zipf_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:90], float(r[vi].replace(',',''))) for r in rows[1:]]
for k,v in seq[-4:]: print('  %9.1f us  %s'%(v/1000.0 if v>5000 else v,k))
PY
echo "== jacobi probe"; timeout 300 python scripts/jacobi_probe.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'block_apply_kernel' -s 40 -c 2 -f -o gpurun_out/r02n_prof_block_apply python scripts/jacobi_probe.py > gpurun_out/r02n_ncu_block_apply.log 2>&1; tail -1 gpurun_out/r02n_ncu_block_apply.log | cut -c1-150
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | grep -E "==|SUMMARY"
