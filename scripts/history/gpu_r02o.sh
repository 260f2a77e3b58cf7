#!/bin/bash
mkdir -p gpurun_out
echo "== pytest csr"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py -q -m gpu --timeout 180 2>&1 | tail -4 | tee gpurun_out/r02o_pytest.txt
echo "== zipf / cfg2"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1; timeout 300 python scripts/zipf_probe.py cfg2 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' --csv --log-file gpurun_out/r02o_zipf_launches.csv python scripts/zipf_probe.py cfg2_zipf > gpurun_out/r02o_zipf_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02o_zipf_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:90], float(r[vi].replace(',',''))) for r in rows[1:]]
for k,v in seq[-4:]: print('  %9.1f us  %s'%(v/1000.0 if v>5000 else v,k))
PY
echo "== bench (spmv only)"; timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu --no-cg --no-gmres > gpurun_out/r02o_bench.json 2>gpurun_out/r02o_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(json.dumps(d['config'].get('twins', d['config']), indent=0)[:1500])"
