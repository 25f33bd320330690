#!/bin/bash
# multi-GPU bench: segmented (three graphs, eager all-reduces between them) vs whole-step graph with the
# NCCL all-reduces captured inside (per-discriminator exchange under the other discriminators' compute)
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02_multi${N}_summary.log
: > $S
run() { local name=$1 secs=$2; shift 2; echo "== $name"; timeout "$secs" "$@" > "$O/r02_multi${N}_$name.log" 2>&1; echo "$name rc=$?" | tee -a $S; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run whole 240 $TR bench.py --gpus $N --steps 20 --warmup 5
if ! grep -q '"metric"' $O/r02_multi${N}_whole.log || [ "${SEG:-0}" = "1" ]; then
  run segmented 240 $TR bench.py --gpus $N --steps 20 --warmup 5 --segmented
fi
if [ "${ONE:-1}" = "1" ]; then run one 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline; fi
for f in whole segmented one; do grep -h '"metric"' $O/r02_multi${N}_$f.log 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$f: n=%d %.2f ms/step  %.1f stories/s' % (d['n_gpus'], d['ms_per_step'], d['value']))
" | tee -a $S; done
