# fragment-layout epilogue (dbg 64) vs SMEM-transposed epilogue (dbg 0), fp16 operands
python tools/layer_bench.py 32 17920 128 128 3 1 1 0 16 0 64
python tools/layer_bench.py 32 17920 128 128 3 1 1 1 both 0 64
python tools/layer_bench.py 32 17920 128 128 11 1 1 1 both 0 64
python tools/layer_bench.py 32 2560 256 256 3 1 1 0 16 0 64
python tools/layer_bench.py 32 2560 256 256 3 1 1 1 both 0 64
python tools/layer_bench.py 32 2560 256 256 11 1 1 1 both 0 64
python tools/layer_bench.py 32 53760 64 64 3 1 1 0 16 0 64
python tools/layer_bench.py 32 53760 64 64 3 1 1 1 both 0 64
python tools/layer_bench.py 1 65536 120 720 1 1 1 0 16 0 64
python tools/layer_bench.py 1 16384 1248 208 1 1 1 1 both 0 64
