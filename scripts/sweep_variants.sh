#!/bin/bash
# usage: scripts/sweep_variants.sh "N:E [opts]" name1 name2 ...   (variants built by build_variants.sh)
cfg=$1; shift
for v in "$@"; do
  echo "== $v"
  NEKCEM_B200_LIB=$PWD/nekcem_b200/lib/variants/$v.so python scripts/sweep.py $cfg 2>&1 | grep -v "l2 err"
done
