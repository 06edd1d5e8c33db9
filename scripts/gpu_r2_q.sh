#!/bin/bash
# ncu full capture of the TMA-fed thin conv (one 4->4 launch of the microbenchmark)
mkdir -p gpurun_out
O=gpurun_out; TAG=r2q
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_thin4t_kernel -s 4 -c 1 -f -o $O/${TAG}_prof_thin python scripts/thin_exp.py > $O/ncu_thin.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_prof_thin.ncu-rep > $O/${TAG}_conv_thin4t_ncu_full.txt 2>&1
python scripts/ncu_hot.py $O/${TAG}_prof_thin.ncu-rep 40 >> $O/${TAG}_conv_thin4t_ncu_full.txt 2>&1
rm -f $O/${TAG}_prof_thin.ncu-rep
tail -2 $O/ncu_thin.log; grep -E "issue_active|pipe_fma|lsu_wave|stalled|dram_throughput|registers_per|gpu__time" $O/${TAG}_conv_thin4t_ncu_full.txt | head -24; grep -A50 "total samples" $O/${TAG}_conv_thin4t_ncu_full.txt | head -50
