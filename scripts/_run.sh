bash scripts/sweep_variants.sh "15:24 const_metrics=0 pipeline=1" w16
bash scripts/sweep_variants.sh "12:26 const_metrics=0 pipeline=1" w13
bash scripts/sweep_variants.sh "11:28 const_metrics=0 pipeline=1" w12
bash scripts/sweep_variants.sh "10:32 const_metrics=0 pipeline=1" w11
