#!/bin/bash
# Multi-GPU call (N = $NG GPUs of one box): tiled single clip with neighbour halo exchange, training step with NCCL
# all-reduce, clip-sharded bench (weak scaling + e2e), 64-clip strong sharding, per-rank D2H bandwidth.
NG=${NG:-2}
mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port"
nvidia-smi topo -m > $O/multi${NG}_topo.txt 2>&1
timeout 600 $T 29611 scripts/bench_tiled.py --frames 10 --h 270 --w 480 > $O/multi${NG}_tiled_4k.json 2> $O/multi${NG}_tiled.err
echo "tiled 4k rc $?"; cat $O/multi${NG}_tiled_4k.json; tail -2 $O/multi${NG}_tiled.err
timeout 600 $T 29612 scripts/bench_tiled.py --frames 10 --h 135 --w 240 > $O/multi${NG}_tiled_1080p.json 2>> $O/multi${NG}_tiled.err
echo "tiled 1080p rc $?"; cat $O/multi${NG}_tiled_1080p.json
timeout 600 $T 29613 scripts/bench_train.py --shape v7 --steps 5 --warmup 4 --graphs > $O/multi${NG}_train_v7.json 2> $O/multi${NG}_train.err
echo "train v7 rc $?"; cat $O/multi${NG}_train_v7.json; tail -2 $O/multi${NG}_train.err
timeout 600 $T 29614 scripts/bench_train.py --shape crop --steps 3 --warmup 4 --graphs > $O/multi${NG}_train_crop.json 2>> $O/multi${NG}_train.err
echo "train crop rc $?"; cat $O/multi${NG}_train_crop.json
timeout 900 $T 29615 bench.py --gpus $NG --steps 5 --warmup 3 > $O/multi${NG}_bench.json 2> $O/multi${NG}_bench.err
echo "bench rc $?"; python -c "
import json
d=json.loads(open('$O/multi${NG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'e2e(u8)', round(d['e2e']['value'],1), 'wall', round(d['e2e']['wall_value'],1), 'e2e_f32', round(d['e2e_f32']['value'],1), 'f32 d2h GB/s/rank', round(d['e2e_f32']['d2h_gbs_per_rank'],1))
"; tail -2 $O/multi${NG}_bench.err
if [ -n "$STRONG" ]; then
timeout 900 $T 29616 bench.py --gpus $NG --total-clips 64 --frames 20 --steps 1 --warmup 1 --no-e2e > $O/multi${NG}_bench_64clips.json 2>> $O/multi${NG}_bench.err
echo "64 clips rc $?"; python -c "
import json
d=json.loads(open('$O/multi${NG}_bench_64clips.json').read().strip().splitlines()[-1])
print('64 clips x 20 frames strong-sharded:', round(d['value'],1), 'fps', d['scaling'], d['config'])
"
fi
if [ -n "$WC" ]; then
CRFP_PIN_WC=1 timeout 900 $T 29617 bench.py --gpus $NG --steps 5 --warmup 3 > $O/multi${NG}_bench_wc.json 2> $O/multi${NG}_bench_wc.err
echo "bench WC rc $?"; python -c "
import json
d=json.loads(open('$O/multi${NG}_bench_wc.json').read().strip().splitlines()[-1])
print('WC: value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'wall', round(d['e2e']['wall_value'],1), d['e2e']['pinned'])
"; tail -2 $O/multi${NG}_bench_wc.err
fi
timeout 120 python scripts/d2h_bw.py > $O/multi${NG}_d2h_single.txt 2>&1; cat $O/multi${NG}_d2h_single.txt
