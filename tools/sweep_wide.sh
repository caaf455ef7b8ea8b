python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f e2e %.3e host_enq %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d.get('host_enqueue_ms_per_step',0)))"; }
for l in 0 4 3 2; do echo "LPT=$l"; GORT_WIDE_LPT=$l run; done
echo "no xcall:"; GORT_NO_XCALL=1 run
