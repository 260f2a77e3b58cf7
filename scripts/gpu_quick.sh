#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -8
