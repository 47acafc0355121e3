#!/bin/bash
for b in 128 256; do
  echo "== subseq $b"
  DALIB200_JPEG_SUBSEQ_BYTES=$b python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(round(d['value']), d['cpu_baseline']['parity_mismatching_elements'], {k: round(v['ms_per_step'],3) for k,v in d['kernels'].items() if 'huff' in k})"
done
