(time python bench.py --steps 5 --warmup 3 --no-cpu-baseline) 2> gpurun_out/bench_r2_c.err | tee gpurun_out/bench_r2_c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
for k,v in d['extra'].items(): print(k, {kk:v[kk] for kk in v if kk in ('value','ms_per_step','roofline_frac','bytes_per_node_stage','l2_norm_of_fields','l2_error_vs_analytic','gpu_launches','setup_s')})
"
tail -4 gpurun_out/bench_r2_c.err
