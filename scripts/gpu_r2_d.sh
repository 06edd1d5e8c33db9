#!/bin/bash
# Round 2, GPU call D (1 GPU): new f3 / tiling / streaming tests, streaming latency, tiled bench on one GPU, training round
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_f3.py tests/test_gpu_tiling.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider > $O/r2d_tests.log 2>&1
echo "tests rc $?"; tail -4 $O/r2d_tests.log
timeout 300 python scripts/bench_stream.py > $O/r2d_stream_1080p.json 2> $O/r2d_stream.err; cat $O/r2d_stream_1080p.json; tail -3 $O/r2d_stream.err
timeout 300 python scripts/bench_stream.py --h 180 --w 320 --frames 40 > $O/r2d_stream_rlit.json 2>> $O/r2d_stream.err; cat $O/r2d_stream_rlit.json
timeout 600 python scripts/bench_tiled.py --frames 10 > $O/r2d_tiled_1gpu_4k.json 2> $O/r2d_tiled.err; cat $O/r2d_tiled_1gpu_4k.json; tail -3 $O/r2d_tiled.err
TAG=r2d bash scripts/gpu_train_round.sh > $O/r2d_train_round.log 2>&1; tail -30 $O/r2d_train_round.log
for f in $O/r2d_train_kernel_times_v7*.txt; do echo "== $f"; head -14 $f; done
cat $O/r2d_train_stock_pytorch_v7.txt | tail -3
