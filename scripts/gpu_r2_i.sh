#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/stream_host_times.py > gpurun_out/r2i_stream_host.txt 2> gpurun_out/r2i_stream.err
timeout 300 python scripts/stream_host_times.py --h 180 --w 320 >> gpurun_out/r2i_stream_host.txt 2>> gpurun_out/r2i_stream.err
cat gpurun_out/r2i_stream_host.txt; tail -5 gpurun_out/r2i_stream.err
