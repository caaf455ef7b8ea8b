run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac']))"; }
for l in 0 0 4 2; do echo "LPT=$l"; GORT_WIDE_LPT=$l run; done
