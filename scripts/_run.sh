bash scripts/sweep_variants.sh "7:48 const_metrics=0 pipeline=1" n256 n320 n288 n256
