# Round-1 evidence run (one B200): GPU test-suite, smoke, both bench arms, launch lists, pair-kernel timelines.
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r01_vocoder_fp16_final.csv $B > gpurun_out/ncu_final.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_r01_encoder_final.csv python tools/encoder_only.py 1024 fp16 > gpurun_out/ncu_enc_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resblock_pair_kernel --launch-skip 63 --launch-count 1 -o gpurun_out/full_r01_fused_s3k3_final -f $B > gpurun_out/ncu_f_final.log 2>&1
python tools/trace_pair.py 32 107520 32 3 1 fp32 > gpurun_out/trace_pair_final.log 2>&1
python tools/trace_pair.py 32 53760 64 7 3 fp32 >> gpurun_out/trace_pair_final.log 2>&1
python tools/trace_pair.py 32 107520 32 3 1 none >> gpurun_out/trace_pair_final.log 2>&1
python tools/pair_bench.py > gpurun_out/pair_bench_final.log 2>&1
python tools/bench_sharded.py --clips 256 > gpurun_out/sharded_n1_final.json 2> gpurun_out/sharded_n1_final.err
cat gpurun_out/pytest_final.log; tail -2 gpurun_out/smoke_final.log
