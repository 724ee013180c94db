timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/sweep_variants.sh "8:36 const_metrics=0 pipeline=1" v9c v9d
bash scripts/sweep_variants.sh "9:32 const_metrics=0 pipeline=1" v10d
python scripts/sweep.py 7:48 8:40 9:36 const_metrics=0,1 pipeline=1
