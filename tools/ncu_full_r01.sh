B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:resblock_pair_kernel --launch-skip 54 --launch-count 1 -o gpurun_out/full_r01_fused_s2k3 -f $B > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resblock_pair_kernel --launch-skip 69 --launch-count 1 -o gpurun_out/full_r01_fused_s3k11 -f $B > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_engine_pair_kernel --launch-skip 133 --launch-count 1 -o gpurun_out/full_r01_pair16_s0k11 -f $B > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_engine_pair_kernel --launch-skip 152 --launch-count 1 -o gpurun_out/full_r01_pair16_s1k11 -f $B > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwconv_tma_kernel --launch-skip 30 --launch-count 1 -o gpurun_out/full_r01_dwconv_tma -f python tools/encoder_only.py 256 fp16 > gpurun_out/ncu_f5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
