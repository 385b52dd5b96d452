mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/t7_tests.log 2>&1
(BFQ_COMMIT_TRACE=1 BFQ_BUILD_TRACE=1 timeout 240 python tools/commit_bench.py > gpurun_out/t7_commit.json 2> gpurun_out/t7_commit.trace.txt)
for C in 3 5 6; do (BFQ_SUBBATCHES=$C timeout 150 python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1) > gpurun_out/t7_sub$C.log 2>&1; done
tail -3 gpurun_out/t7_tests.log; cut -c1-600 gpurun_out/t7_commit.json; grep "bfq build" gpurun_out/t7_commit.trace.txt | head -8
for C in 3 5 6; do python - <<PY
import json
l=json.loads(open('gpurun_out/t7_sub$C.log').read().strip().splitlines()[-1])
print($C, l['value'], l['e2e']['value'], l['e2e'].get('last_step_breakdown_ms'))
PY
done
