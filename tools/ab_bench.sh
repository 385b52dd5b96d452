#!/bin/bash
# A/B the bench over builds / switches: tools/ab_bench.sh "BFQ_LIB=path/to/lib.so" "BFQ_ORDER=0" ... (one run per argument;
# an argument is a space-separated list of VAR=VALUE pairs)
for envs in "$@"; do
  echo "== $envs"
  env $envs python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.4f kernel_ms %.4f frac %.4f e2e %.4g launches %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']))"
done
