#!/bin/bash
# torchrun --no-python wrapper: the rank named by NCU_RANK runs under ncu (kernel durations only), the others plain
if [ "$RANK" == "${NCU_RANK:-1}" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' \
       --csv --log-file gpurun_out/${NCU_TAG:-rank}_launches_rank$RANK.csv python "$@"
else
  exec python "$@"
fi
