#!/bin/bash
# A/B the bench over several builds of the library: tools/ab_bench.sh lib1.so lib2.so ... (BFQ_LIB override)
for lib in "$@"; do
  echo "== $lib"
  BFQ_LIB=$lib python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g ms/step %.4f kernel_ms %.4f frac %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))"
done
