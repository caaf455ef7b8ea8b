# in-kernel timeline of rsurf_wide_kernel in the pipelined steady state (call 30 of a bench run)
GORT_TIMELINE=30 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>&1 | grep -v '^{'
