set -x
# stage 1 conv1 k3 (fp16 in, fp16 out), conv2 (res, both)
python tools/layer_bench.py 32 17920 128 128 3 1 1 0 16 0 1 2 4 8 9 15
python tools/layer_bench.py 32 17920 128 128 3 1 1 1 both 0 1 8 9
python tools/layer_bench.py 32 17920 128 128 11 1 1 0 16 0 1 8
# stage 0 k3
python tools/layer_bench.py 32 2560 256 256 3 1 1 0 16 0 1 8
python tools/layer_bench.py 32 2560 256 256 11 1 1 0 16 0 1 8
# stage 2 / 3
python tools/layer_bench.py 32 53760 64 64 3 1 1 0 16 0 1 8 9
python tools/layer_bench.py 32 53760 64 64 11 1 1 0 16 0 1 8 9
python tools/layer_bench.py 32 107520 32 32 3 1 1 0 16 0 1 8 9
python tools/layer_bench.py 32 107520 32 32 11 1 1 0 16 0 1 8 9
python tools/layer_bench.py 32 107520 32 32 11 1 1 1 both 0 1 8 9
