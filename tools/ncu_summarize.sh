#!/bin/bash
# usage: tools/ncu_summarize.sh gpurun_out/X.ncu-rep profiles/<name>   -> <name>_raw.txt (selected counters per kernel),
# <name>_sass_<kernel>.txt (SASS opcode mix + top stall lines).  Run in the build container (no GPU needed).
set -e
rep=$1; out=$2
here=$(dirname "$0")
ncu -i "$rep" --page raw --csv 2>/dev/null | python "$here/ncu_raw.py" 'smsp__average_warps_issue_stalled.*ratio|lts__t_sector_hit_rate|dram__throughput|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$' > "${out}_raw.txt"
shift 2
for k in "$@"; do
  ncu -i "$rep" --page source --csv --kernel-name "regex:$k" 2>/dev/null | python "$here/ncu_sass_mix.py" 30 > "${out}_sass_${k}.txt" || true
done
