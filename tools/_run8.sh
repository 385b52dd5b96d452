mkdir -p gpurun_out
(timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/t8_tests.log 2>&1
(BFQ_COMMIT_TRACE=1 BFQ_BUILD_TRACE=1 timeout 240 python tools/commit_bench.py > gpurun_out/t8_commit.json 2> gpurun_out/t8_commit.trace.txt)
(BFQ_UPLOAD=plain BFQ_COMMIT_TRACE=1 timeout 240 python tools/commit_bench.py 2>&1 | grep "full commit" | head -8) > gpurun_out/t8_commit_plain.txt 2>&1
tail -3 gpurun_out/t8_tests.log; cut -c1-700 gpurun_out/t8_commit.json; grep "bfq build\|full commit" gpurun_out/t8_commit.trace.txt | head -14; echo PLAIN; cat gpurun_out/t8_commit_plain.txt
