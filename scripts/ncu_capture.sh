#!/bin/bash
# usage: scripts/ncu_capture.sh "N:E" tag [sweep options...]  -- full ncu capture of one stage-kernel launch
cfg=$1; tag=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:'slab_kernel|pipe_kernel|stage_kernel' -s 12 -c 1 \
    -f -o gpurun_out/prof_$tag python scripts/sweep.py $cfg:1 "$@" > gpurun_out/prof_$tag.log 2>&1
