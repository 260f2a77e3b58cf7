#!/bin/bash
# r02l: tensor-core many-rhs block apply (tests + table), per-kernel breakdown of the Zipf SpMV
mkdir -p gpurun_out
echo "== pytest many rhs"; timeout 900 python -m pytest tests/test_jacobi_adaptive_gpu.py -q -m gpu --timeout 180 -k "many_rhs" 2>&1 | tail -6 | tee gpurun_out/r02l_pytest_mma.txt
echo "== pytest jacobi parity (plain)"; timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_solvers_gpu.py -q -m gpu --timeout 180 -k "jacobi or Jacobi or gmres" 2>&1 | tail -3
echo "== zipf"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' --csv --log-file gpurun_out/r02l_zipf_launches.csv python scripts/zipf_probe.py cfg2_zipf > gpurun_out/r02l_zipf_ncu.log 2>&1; tail -1 gpurun_out/r02l_zipf_ncu.log
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02l_zipf_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
last=collections.OrderedDict()
seq=[(r[ki][:90], float(r[vi].replace(',',''))) for r in rows[1:]]
print(len(seq),'launches; last 12:')
for k,v in seq[-12:]: print('  %9.1f us  %s'%(v/1000.0 if v>5000 else v,k))
PY
echo "== kernel table"; EXP_KERNELS_OUT=gpurun_out/r02l_kernels_roofline.json timeout 900 python scripts/exp_kernels.py > gpurun_out/r02l_exp_kernels.log 2>&1; tail -14 gpurun_out/r02l_exp_kernels.log | cut -c1-230
