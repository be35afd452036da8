# encoder IR-stage GEMM shapes, fp16: production dispatch (pair kernel) with probe switches
python tools/layer_bench.py 1 65536 120 720 1 1 1 0 16 0 1 8 9
python tools/layer_bench.py 1 65536 720 120 1 1 1 1 both 0 1 8 9
python tools/layer_bench.py 1 16384 208 1248 1 1 1 0 16 0 1 8 9
python tools/layer_bench.py 1 16384 1248 208 1 1 1 1 both 0 1 8 9
