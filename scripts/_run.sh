(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/b8.err) 2> gpurun_out/b8.time | tee gpurun_out/bench_r2_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'], d['e2e']['host_buffers'], d['config']['face_exchange'])
for k,v in d['extra'].items(): print(k, {kk:v[kk] for kk in v if kk in ('value','ms_per_step','roofline_frac','bitwise_equal','max_abs_diff','mesh')})
"
tail -3 gpurun_out/b8.err; cat gpurun_out/b8.time; free -g | head -2
