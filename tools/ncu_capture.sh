#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <skip> <count> <python script + args...>
# Captures --set full for a few launches and exports the raw/source pages as CSV (small files).
name=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count \
    -f -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/$name.details.txt 2>/dev/null
ls -la gpurun_out/$name.*
sz=$(stat -c %s gpurun_out/$name.ncu-rep)
if [ "$sz" -gt 20000000 ]; then rm -f gpurun_out/$name.ncu-rep; fi
tail -3 gpurun_out/$name.log
