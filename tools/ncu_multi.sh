#!/bin/bash
# usage: tools/ncu_multi.sh <tag> <kernel-regex> <launch-skip> <launch-count> <python args...>   (run on the GPU box)
# `ncu --set full` over the launches matching the regex; raw metrics only (gzipped CSV) under gpurun_out/.
TAG=$1; KRE=$2; SKIP=$3; CNT=$4; shift 4
timeout 500 ncu --set full --clock-control none -k "regex:$KRE" -s $SKIP -c $CNT -f -o /tmp/$TAG "$@" > gpurun_out/$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/$TAG.raw.csv.gz
ls -la gpurun_out/$TAG.raw.csv.gz | awk '{print $5, $9}'
