#!/bin/bash
# spill stores/loads of one kernel instantiation per barrier-delimited segment: tools/exp/spills.sh decode IfLi10ELb1
cd /tmp && rm -f /tmp/*.cubin && cuobjdump -xelf all /tmp/$1.o >/dev/null 2>&1
nvdisasm -g -c /tmp/$1.sm_100a.cubin 2>/dev/null | awk -v pat="$2" '/^\.text\./ {f = ($0 ~ pat)} f' > /tmp/$1_lines.sass
grep -n -E "STL|LDL|REDG|CCTL.IVALL|BAR.SYNC" /tmp/$1_lines.sass | awk -F: '{print $1, $2}' | awk '{k=($0 ~ /STL/)?"STL":(($0 ~ /LDL/)?"LDL":(($0 ~ /REDG/)?"ARRIVE":(($0 ~ /IVALL/)?"WAIT":"BAR"))); print $1, k}' | awk '{if ($2=="STL"||$2=="LDL") {c[$2]++} else {if (c["STL"]+c["LDL"]>0) printf "%s@%s  (since prev: STL %d LDL %d)\n", $2, $1, c["STL"], c["LDL"]; c["STL"]=0; c["LDL"]=0}} END {printf "END (STL %d LDL %d)\n", c["STL"], c["LDL"]}'
wc -l /tmp/$1_lines.sass
