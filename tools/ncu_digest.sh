#!/usr/bin/env bash
# Digest of an .ncu-rep: the key raw metrics used in profiles/*.md
for f in "$@"; do
  echo "== $f"
  ncu -i "$f" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
r=list(csv.reader(sys.stdin))
hdr=r[0]; units=r[1]; 
keys=["Kernel Name","Grid Size","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","smsp__cycles_active.avg","sm__cycles_elapsed.max","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed.sum"]
for row in r[2:]:
    for k in keys:
        for i,h in enumerate(hdr):
            if h==k: print("  %-70s %s %s"%(k,row[i],units[i]))
'
done
