# A/B of the BRDF pipeline switches on the C2 bench (run on a B200): default, per-thread stores instead of TMA
# bulk stores, no cross-call overlap, no programmatic dependent launch at all
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f isolated kernel ms %.4f frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac']))"; }
echo "default"; run
echo "GORT_NO_TMA=1"; GORT_NO_TMA=1 run
echo "GORT_NO_XCALL=1"; GORT_NO_XCALL=1 run
echo "GORT_NO_PDL=1"; GORT_NO_PDL=1 run
