#!/bin/bash
# Round-2 evidence, one GPU (run through gpurun): bench lines, ncu launch list, ncu --set full captures, parity report.
# Everything lands in gpurun_out/r2/ (kept under 64 MiB: the big capture is reduced to its raw CSV on the box).
mkdir -p gpurun_out/r2; rm -f gpurun_out/*.ncu-rep
(timeout 600 python bench.py > gpurun_out/r2/r2_bench.json 2> gpurun_out/r2/r2_bench.err; echo bench rc=$?)
(timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2/r2_bench_20.json 2>> gpurun_out/r2/r2_bench.err; echo bench20 rc=$?)
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/r2_bench_reference.json 2>> gpurun_out/r2/r2_bench.err; echo reference rc=$?)
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /tmp/ncu_launch.log 2>&1; echo launches rc=$?)
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:"rsurf_wide|geom_kernel" -s 8 -c 2 -o gpurun_out/r2/r2_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > /tmp/ncu_bench.log 2>&1; echo ncu-bench rc=$?)
(timeout 900 ncu --set full --clock-control none -o /tmp/r2_all python tools/prof_all.py > /tmp/ncu_all.log 2>&1; echo ncu-all rc=$?; ncu -i /tmp/r2_all.ncu-rep --page raw --csv > gpurun_out/r2/r2_all_raw.csv 2>/dev/null)
(timeout 600 ncu --set full --import-source on --clock-control none -k regex:"lut_tube|lut_crown_kernel<1>|lut_crown_kernelILi1" -c 2 -o gpurun_out/r2/r2_lut python tools/dev_lut_ncu.py 8192 > /tmp/ncu_lut.log 2>&1; echo ncu-lut rc=$?)
(GORT_TIMELINE=30 timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2> gpurun_out/r2/r2_timeline.txt > /dev/null; echo timeline rc=$?)
(timeout 600 python tools/parity_report.py --out gpurun_out/r2/r2_parity.json > gpurun_out/r2/r2_parity_headline.json 2> gpurun_out/r2/r2_parity.err; echo parity rc=$?)
ls -la gpurun_out/r2; du -sh gpurun_out
