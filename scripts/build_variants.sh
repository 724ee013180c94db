#!/bin/bash
# Developer tool: build libnekcem_b200 variants with different -D tunables of one stage kernel file.
# usage: scripts/build_variants.sh [pipe|slab] name1="-DPIPE_ONLY_N=16 -DPIPE_EPI=8" name2="..." ...
# Each variant lands in nekcem_b200/lib/variants/<name>.so; select it with NEKCEM_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../nekcem_b200/csrc"
which=pipe
if [ "$1" = pipe ] || [ "$1" = slab ]; then which=$1; shift; fi
other=$([ $which = pipe ] && echo slab || echo pipe)
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin /usr/bin/g++"
extra=""; [ $which = pipe ] && extra="-DPIPE_SINGLE"
make -s _obj/nekcem_b200.o _obj/stage2d.o _obj/stage2d_strict.o _obj/graphene.o _obj/graphene_strict.o _obj/fortran_abi.o _obj/stage_slab.o _obj/stage_slab_strict.o _obj/stage_slab_hi.o _obj/stage_slab_hi_strict.o _obj/stage_sweep.o
[ $which = slab ] && make -s _obj/stage_pipe_dispatch.o $(ls _obj/stage_pipe_n*.o 2>/dev/null)
pipeobjs=$([ $which = slab ] && ls _obj/stage_pipe_dispatch.o _obj/stage_pipe_n*.o || echo _obj/stage_slab.o)
mkdir -p ../lib/variants _obj/variants
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  (
    $NVCC $FLAGS $extra $defs -Xptxas -v -c stage_$which.cu -o _obj/variants/$name.o 2> _obj/variants/$name.log
    $NVCC $ARCH -shared -o ../lib/variants/$name.so _obj/nekcem_b200.o _obj/stage2d.o _obj/stage2d_strict.o _obj/graphene.o _obj/graphene_strict.o _obj/variants/$name.o $pipeobjs _obj/stage_slab_strict.o _obj/stage_slab_hi.o _obj/stage_slab_hi_strict.o _obj/stage_sweep.o _obj/fortran_abi.o -lcudart -ldl
    grep -E "Used|spill" _obj/variants/$name.log | paste - - | awk -v n=$name '{print n": "$0}' | sed 's/ptxas info    ://g' | head -4
  ) &
done
wait
