#!/bin/bash
# Run ON THE GPU BOX after ncu: turn every gpurun_out/*.ncu-rep into small CSVs (raw metrics +
# per-instruction source page, gzipped) and delete the report (gpurun_out is capped at 64 MiB).
for rep in gpurun_out/*.ncu-rep; do
  [ -e "$rep" ] || continue
  base="${rep%.ncu-rep}"
  ncu -i "$rep" --page raw --csv > "$base.raw.csv" 2>/dev/null
  ncu -i "$rep" --page source --csv --print-source sass 2>/dev/null | gzip -9 > "$base.src.csv.gz"
  rm -f "$rep"
done
ls -la gpurun_out
