#!/bin/bash
# usage: tools/ncu_capture.sh <tag> <kernel-regex> <python args...>   (run on the GPU box)
# One `ncu --set full` capture of the first matching launch; exports the raw metrics and the source/SASS page as
# gzipped CSV under gpurun_out/ and removes the .ncu-rep (the reports exceed gpurun's 64 MiB copy-back limit).
TAG=$1; KRE=$2; shift 2
REP=/tmp/$TAG.ncu-rep
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$KRE" -c 1 -f -o /tmp/$TAG "$@" > gpurun_out/$TAG.log 2>&1
ncu -i $REP --page raw --csv 2>/dev/null | gzip > gpurun_out/$TAG.raw.csv.gz
ncu -i $REP --page source --csv 2>/dev/null | gzip > gpurun_out/$TAG.src.csv.gz
ls -la $REP gpurun_out/$TAG.* | awk '{print $5, $9}'
