python scripts/sweep.py 3:56 5:48 6:48 7:48 8:40 9:36 10:32 11:30 12:28 13:26 14:24 15:24 const_metrics=0,1 > gpurun_out/r2_sweep_all_orders.txt 2>&1
tail -4 gpurun_out/r2_sweep_all_orders.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2_launches_bench_n7_e64.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
grep -c pipe_kernel gpurun_out/r2_launches_bench_n7_e64.csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
