# Round-2 evidence run (one B200): GPU test-suite with its printed numbers, smoke, both bench arms, encoder / vocoder
# launch lists (ncu, serialised), one `ncu --set full` capture of the encoder's top kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -rA 2>&1 | grep -E "^\[|scaled|cuDNN|rel err|SNR|passed|failed|PASSED|FAILED" > gpurun_out/pytest_gpu_r02_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_r02_final.log
python bench.py > gpurun_out/bench_r02_n1_final.json 2> gpurun_out/bench_r02_n1_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_n1_reference.json 2> gpurun_out/bench_r02_n1_reference.err
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
python tools/enc_once.py 2048 55 > gpurun_out/plain_enc.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_encoder_final.csv python tools/enc_once.py 2048 55 > gpurun_out/ncu_enc_final.log 2>&1
B="python bench.py --workload vocoder --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain_voc.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_vocoder_final.csv $B > gpurun_out/ncu_voc_final.log 2>&1
python tools/enc_once.py 1024 55 > gpurun_out/plain_enc2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_expand_dw_kernel -s 8 -c 1 -o gpurun_out/prof_r02_expand_dw_final -f python tools/enc_once.py 1024 55 > gpurun_out/ncu_full_final.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02_final.log; tail -2 gpurun_out/smoke_r02_final.log; head -c 300 gpurun_out/bench_r02_n1_final.json
