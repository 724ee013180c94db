timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/sweep.py 3:56 5:48 6:48 7:48 8:40 9:36 10:32 11:30 12:28 13:26 14:24 15:24 const_metrics=0,1 > gpurun_out/r2_sweep_all_orders.txt 2>&1
grep -c "N=" gpurun_out/r2_sweep_all_orders.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
