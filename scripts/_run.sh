timeout 600 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_parity.py -q -x -k "pml or dielectric" 2>&1 | tail -3
python scripts/_aux.py
