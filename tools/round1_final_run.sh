python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r01_vocoder_fp16_final.csv $B > gpurun_out/ncu_final.log 2>&1
python tools/trace_pair.py 32 107520 32 3 1 fp32 > gpurun_out/trace_pair_final.log 2>&1
python tools/trace_pair.py 32 53760 64 7 3 fp32 >> gpurun_out/trace_pair_final.log 2>&1
python tools/trace_pair.py 32 107520 32 3 1 split >> gpurun_out/trace_pair_final.log 2>&1
python tools/pair_bench.py > gpurun_out/pair_bench_final.log 2>&1
python tools/bench_sharded.py --clips 256 > gpurun_out/sharded_n1_final.json 2> gpurun_out/sharded_n1_final.err
cat gpurun_out/pytest_final.log
