timeout 900 python -m pytest tests/test_gpu_zgraphene.py tests/test_gpu_zz_tworank_local.py -q -x 2>&1 | tail -4
