python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import json,sys
txt=sys.stdin.read().strip().splitlines()
for t in txt[:-1]: print(t)
d=json.loads(txt[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f e2e %.3e host_enq %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d.get('host_enqueue_ms_per_step',0)))"; }
export GORT_NO_TMA=1
for l in 0 4 3 2; do echo "== LPT=$l"; GORT_WIDE_LPT=$l run; done
unset GORT_NO_TMA
for l in 0 3; do echo "== TMA LPT=$l"; GORT_WIDE_LPT=$l run; done
export GORT_NO_TMA=1
echo "== timeline LPT=0"; GORT_TIMELINE=30 run
