python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in "4 3" "4 2" "2 3" "2 2" "2 4"; do set -- $cfg; echo "LPT=$1 MINB=$2"; GORT_WIDE_LPT=$1 GORT_WIDE_MINB=$2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac'],d['e2e']['value']))"; done
echo "no PDL:"; GORT_NO_PDL=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f geom_ms %.4f frac %.3f e2e %.3e'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['geom_kernel_ms'],d['roofline']['frac'],d['e2e']['value']))"
