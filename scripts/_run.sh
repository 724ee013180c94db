NEKCEM_B200_LIB=$PWD/nekcem_b200/lib/variants/g296.so timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "every_order and 16" 2>&1 | tail -1
bash scripts/sweep_variants.sh "15:24 const_metrics=0" g0 g296 g148 g592
bash scripts/sweep_variants.sh "12:26 const_metrics=0" h0 h296
