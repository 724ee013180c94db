bash scripts/sweep_variants.sh "15:24 const_metrics=0" c16 a16 b16
bash scripts/sweep_variants.sh "12:26 const_metrics=0" c13 a13
