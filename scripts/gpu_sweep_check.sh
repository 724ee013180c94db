mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweep.py -q --timeout 180 -p no:cacheprovider > gpurun_out/sweep_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/sweep_tests.log
tail -8 gpurun_out/sweep_tests.log
for cfg in ${SWEEP_CFGS:-15:24 14:24 13:26 12:26 11:30 10:32}; do
  timeout 300 python scripts/sweep.py $cfg const_metrics=0 sweep=0,1 2>&1 | tee -a gpurun_out/sweep_perf.log
done
