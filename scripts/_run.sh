timeout 600 python -m pytest tests/test_gpu_pipe.py -q -x -k "streamed or agrees" 2>&1 | tail -5
(time python bench.py --steps 10 --warmup 3) 2> gpurun_out/bench_r2_a.err | tee gpurun_out/bench_r2_a.json | cut -c1-3000
tail -5 gpurun_out/bench_r2_a.err
