#!/bin/bash
# usage: scripts/ncu_capture.sh "N:E" tag [sweep options...]  -- full ncu capture of one stage-kernel
# launch, summarised on the spot (gpurun_out/prof_<tag>.txt); the report itself is kept only with KEEP_REP=1
cfg=$1; tag=$2; shift 2
ncu --set full --clock-control none --import-source on -k regex:'slab_kernel|pipe_kernel|sweep_kernel|stage2d_kernel' -s 12 -c 1 \
    -f -o gpurun_out/prof_$tag python scripts/sweep.py $cfg:1 "$@" > gpurun_out/prof_$tag.log 2>&1
python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep > gpurun_out/prof_$tag.txt 2>&1
python scripts/ncu_lines.py gpurun_out/prof_$tag.ncu-rep 40 >> gpurun_out/prof_$tag.txt 2>&1
python scripts/ncu_lines.py gpurun_out/prof_$tag.ncu-rep 25 smem >> gpurun_out/prof_$tag.txt 2>&1
[ -n "$KEEP_REP" ] || rm -f gpurun_out/prof_$tag.ncu-rep
