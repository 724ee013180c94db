bash scripts/sweep_variants.sh "7:48 const_metrics=0 pipeline=1" d k128 k32 k160
bash scripts/sweep_variants.sh "7:48 const_metrics=0 pipeline=1 pre:xtrace=0" d k32
