M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active"
B="python bench.py --workload vocoder --steps 1 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/plain_voc.log 2>&1 && \
ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r02_vocoder_res16.csv $B > gpurun_out/ncu_voc_res16.log 2>&1
python bench.py > gpurun_out/bench_r02_n1_res16.json 2> gpurun_out/bench_r02_n1_res16.err
head -c 300 gpurun_out/bench_r02_n1_res16.json
