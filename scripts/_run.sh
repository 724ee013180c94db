timeout 600 bash scripts/ncu_capture.sh 7:40 r2_pipe_n8_allpml const_metrics=0 case:pml=all
timeout 600 bash scripts/ncu_capture.sh 7:512 r2_stage2d_te_n8 case:dim=2
tail -3 gpurun_out/prof_r2_stage2d_te_n8.log
