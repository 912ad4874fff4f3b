#!/bin/bash
# round 2, call Q: ncu source-level profile of the fused residual-unit kernel (C = 64, 128) and the lo-accumulator conv kernel (C = 256 k7);
# reports are exported to CSV on the box (the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out/r02q /tmp/ncu
for spec in "c64:26" "c128:30" "c256k7:34"; do
  tag=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_ru_fused|conv_umma" --launch-skip $skip --launch-count 1 -o /tmp/ncu/$tag python scripts/one_forward.py 4 10 reps=2 > gpurun_out/r02q/ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
  ncu -i /tmp/ncu/$tag.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r02q/src_$tag.csv.gz
  ncu -i /tmp/ncu/$tag.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02q/raw_$tag.csv
  ncu -i /tmp/ncu/$tag.ncu-rep --page details 2>/dev/null | head -200 > gpurun_out/r02q/details_$tag.txt
done
ls -la gpurun_out/r02q
