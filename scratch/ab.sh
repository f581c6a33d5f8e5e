#!/bin/bash
# A/B harness (not product code): for every scratch/variants/*.so, time the late-regime kernel on c3_small and the
# full C3 bench.  Usage: scratch/ab.sh <outfile> [variants...]
out=$1; shift
python - <<'PY' >> $out 2>&1
import torch
a=torch.empty(1<<28,dtype=torch.float32,device='cuda'); b=torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): b.copy_(a)
e1.record(); torch.cuda.synchronize()
print('copy GB/s', 10*2*a.numel()*4/e0.elapsed_time(e1)/1e6)
PY
cp dualip_b200/_lib/libdualip_b200.so /tmp/orig.so
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}; [ "$envs" == "$v" ] && envs=""
  cp scratch/variants/$name.so dualip_b200/_lib/libdualip_b200.so
  for k in mixed simplex box; do env $envs timeout 120 python scratch/exp_lam.py $k 100 2>&1 | grep iter | sed "s/^/$v /" >> $out; done
  if [ -z "$SKIP_FULL" ]; then env $envs timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v C3 it/s %.1f frac %.3f kmin %.3f kmax %.3f clocks %s' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms_min'], d['roofline']['kernel_ms_max'], d['clocks']))" >> $out; fi
done
cp /tmp/orig.so dualip_b200/_lib/libdualip_b200.so
cat $out
