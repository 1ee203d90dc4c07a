#!/bin/bash
# scratch: sweep resident CTAs/SM for knit_outer on the syc-32 workload
cd "$(dirname "$0")/.." && for n in 2 3 4 6 8; do
  QCK_KO_CTAS_PER_SM=$n python bench.py --profile --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ctas/sm', $n, 'ms', round(d['ms_per_step'],3), 'knit ms', round(d['roofline']['kernel_ms'],3), 'GB/s', round(d['roofline']['achieved'],1))"
done
