#!/bin/bash
# compute-sanitizer over the shipped CSR kernels on small matrices (scripts/lab/spmv_lab san): racecheck (shared-memory
# hazards: stage ring, dot epilogue, reductions), synccheck (barrier / mbarrier misuse), memcheck (out-of-bounds).
# Output -> gpurun_out/sanitize_<tool>.log (summaries are committed under profiles/)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 scripts/lab/spmv_lab san > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "san:|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_$tool.log | head -12
done
