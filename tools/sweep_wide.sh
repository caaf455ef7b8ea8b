python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f e2e %.3e host_enq %.4f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac'],d['e2e']['value'],d.get('host_enqueue_ms_per_step',0)))"; }
for cfg in "4 2" "2 2" "2 4" "4 3"; do set -- $cfg; echo "LPT=$1 MINB=$2"; GORT_WIDE_LPT=$1 GORT_WIDE_MINB=$2 run; done
echo "no xcall:"; GORT_NO_XCALL=1 run
echo "no PDL:"; GORT_NO_PDL=1 run
