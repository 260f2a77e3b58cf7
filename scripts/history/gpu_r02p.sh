#!/bin/bash
mkdir -p gpurun_out
echo "== zipf / cfg2"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1; timeout 300 python scripts/zipf_probe.py cfg2 2>&1 | tail -1
echo "== jacobi probe"; timeout 300 python scripts/jacobi_probe.py 2>&1 | tail -4 | tee gpurun_out/r02p_jacobi_probe.txt
echo "== ncu full: zipf main kernel (part 0)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'warp_stream_kernel' -s 12 -c 1 -f -o gpurun_out/r02p_prof_zipf_stream python scripts/zipf_probe.py cfg2_zipf > gpurun_out/r02p_ncu_zipf.log 2>&1; tail -1 gpurun_out/r02p_ncu_zipf.log | cut -c1-150
echo "== pytest"; timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -3 | tee gpurun_out/r02p_pytest.txt
