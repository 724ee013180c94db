timeout 900 python -m pytest tests/test_gpu_pipe.py -q -x -k "3dboxpml or restart" 2>&1 | tail -4
