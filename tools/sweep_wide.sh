run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('  value %.3e ms/step %.4f rsurf_ms %.4f frac %.3f checksum %.12f'%(d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['roofline']['frac'],d['checksum']))"; }
for cfg in "3 192" "4 192" "4 224" "6 192" "3 224" "4 160"; do set -- $cfg; echo "TMAB=$1 LPT4 pick$2"; GORT_TMAB=$1 GORT_WIDE_LPT=4 GORT_WIDE_PICK=$2 run; done
