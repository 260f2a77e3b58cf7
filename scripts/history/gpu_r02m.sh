#!/bin/bash
# r02m: Zipf after (split threshold 1024, warp-summed wide rows), full gpu tests, kernel table (ELL / SELL-P batch 4,
# multi_dot one wave, adaptive apply with the hoisted switch)
mkdir -p gpurun_out
echo "== pytest all"; timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -8 | tee gpurun_out/r02m_pytest.txt
echo "== zipf"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' --csv --log-file gpurun_out/r02m_zipf_launches.csv python scripts/zipf_probe.py cfg2_zipf > gpurun_out/r02m_zipf_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02m_zipf_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:90], float(r[vi].replace(',',''))) for r in rows[1:]]
for k,v in seq[-4:]: print('  %9.1f us  %s'%(v/1000.0 if v>5000 else v,k))
PY
echo "== kernel table"; EXP_KERNELS_OUT=gpurun_out/r02m_kernels_roofline.json timeout 900 python scripts/exp_kernels.py > gpurun_out/r02m_exp_kernels.log 2>&1; grep -E "ell|sellp|hybrid|zipf|multi_dot|simple_apply 250k x 16x16 fp32|adaptive" gpurun_out/r02m_exp_kernels.log | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-8s %-85s %8.4f ms %5.1f%%'%(d['row'],d['kernel'][:85],d['ms'],100*d['frac']))"
echo "== sanitizer (wide rows + threshold)"; bash scripts/gpu_sanitize.sh 2>&1 | grep -E "==|SUMMARY"
