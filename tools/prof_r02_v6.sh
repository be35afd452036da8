M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
python tools/enc_once.py 2048 247 > gpurun_out/plain_enc_v6.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_encoder_v6.csv python tools/enc_once.py 2048 247 > gpurun_out/ncu_enc_v6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mb_project_kernel -s 6 -c 1 -o gpurun_out/prof_r02_project_s4 -f python tools/enc_once.py 1024 247 > gpurun_out/ncu_full_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mb_project_kernel -s 12 -c 1 -o gpurun_out/prof_r02_project_s5 -f python tools/enc_once.py 1024 247 > gpurun_out/ncu_full_p5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_engine_kernel -s 0 -c 1 -o gpurun_out/prof_r02_stage0_b0 -f python tools/enc_once.py 1024 247 > gpurun_out/ncu_full_s0.log 2>&1
tail -2 gpurun_out/ncu_full_p4.log
