#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcn_align_fused -s 3 -c 1 -f -o $O/r2g_prof_fused python scripts/fused_trace.py > $O/r2g_ncu.log 2>&1
tail -3 $O/r2g_ncu.log
python scripts/ncu_summary.py $O/r2g_prof_fused.ncu-rep > $O/r2g_dcn_align_fused_ncu_full.txt 2>&1
python scripts/ncu_hot.py $O/r2g_prof_fused.ncu-rep 60 >> $O/r2g_dcn_align_fused_ncu_full.txt 2>&1
rm -f $O/r2g_prof_fused.ncu-rep
cat $O/r2g_dcn_align_fused_ncu_full.txt | cut -c1-170
