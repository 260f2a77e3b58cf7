#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -8
echo "== exp kernels"; timeout 600 python scripts/exp_kernels.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception:
        print(l.rstrip()[:200]); continue
    print('%-8s %-62s %8.3f ms %7.0f GB/s %5.1f%%' % (d['row'], d['kernel'][:62], d['ms'], d['gbs'], 100*d['frac']))"
