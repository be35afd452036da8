# Round-2 late evidence run (one B200): GPU test-suite with its printed numbers, smoke, both bench arms, encoder / vocoder
# launch lists (ncu, serialised), BiLSTM timings, ncu --set full captures of mb_project / fused_er (stride-2) / lstm cluster.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -rA 2>&1 | grep -E "^\[|scaled|cuDNN|rel err|SNR|passed|failed|PASSED|FAILED" > gpurun_out/pytest_gpu_r02_late.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_late.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_late.log
python bench.py > gpurun_out/bench_r02_n1_late.json 2> gpurun_out/bench_r02_n1_late.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_n1_reference_late.json 2> gpurun_out/bench_r02_n1_reference_late.err
python tools/lstm_time.py 64 600 fp16 > gpurun_out/lstm_time_r02_late.log 2>&1
python tools/lstm_time.py 1 150 fp16 >> gpurun_out/lstm_time_r02_late.log 2>&1
M2S_LSTM_CLUSTER=0 python tools/lstm_time.py 64 600 fp16 >> gpurun_out/lstm_time_r02_late.log 2>&1
python tools/enc_time.py 4096 0 55 119 247 > gpurun_out/enc_time_r02_late.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
python tools/enc_once.py 2048 247 > gpurun_out/plain_enc.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_encoder_late.csv python tools/enc_once.py 2048 247 > gpurun_out/ncu_enc_late.log 2>&1
B="python bench.py --workload vocoder --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain_voc.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_vocoder_late.csv $B > gpurun_out/ncu_voc_late.log 2>&1
python tools/enc_once.py 1024 247 > gpurun_out/plain_enc2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_project_kernel -s 6 -c 1 -o gpurun_out/prof_r02_project_s4_late -f python tools/enc_once.py 1024 247 > gpurun_out/ncu_full_late_p4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_er_kernel -s 0 -c 1 -o gpurun_out/prof_r02_fused_er_s2d_late -f python tools/enc_once.py 1024 247 > gpurun_out/ncu_full_late_er.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_cluster_kernel -s 2 -c 1 -o gpurun_out/prof_r02_lstm_cluster_late -f python tools/lstm_time.py 64 600 fp16 > gpurun_out/ncu_full_late_lstm.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02_late.log; tail -2 gpurun_out/smoke_r02_late.log; head -c 300 gpurun_out/bench_r02_n1_late.json
