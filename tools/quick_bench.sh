#!/bin/bash
# step / e2e / per-kernel device times of the bench configurations (no CPU baseline)
for c in ${@:-C5 C2 C4}; do
  python bench.py --config $c --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); k=d['roofline']['kernel_ms']
print('$c', 'step %.1f us'%(d['ms_per_step']*1e3), 'e2e %.1f us'%(d['e2e']['ms_per_step']*1e3), 'alone: pass %.1f foreign %.1f epilogue %.1f'%(k['pass_kernel']*1e3,k['foreign_kernel']*1e3,k['epilogue_kernel']*1e3))"
done
