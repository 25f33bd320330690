#!/bin/bash
# count deviating step losses of tools/diag_cascade_graph.py over several processes for one flag setting
name=$1; shift
for i in 1 2 3 4 5; do timeout 200 python tools/diag_cascade_graph.py $name 2>&1 | grep "^$name" | awk '{print $3,$6,$9}' ; done | sort | uniq -c | sort -k2,2n -k1,1rn
