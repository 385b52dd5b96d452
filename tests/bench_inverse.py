#!/usr/bin/env python
"""Record-keeping run of the inverse path on BASELINE.json config C5 (1M retained topics vs 100k wildcard SUBSCRIBE
filters): filters/s through bfq_rmatch (host buffers in, ids out) with limit = 10 (RetainMessageMatchLimit default) and
unlimited, next to the oracle's TopicLevelTrie restatement on the host cores. Not the driver's bench line (bench.py); it lives
under tests/ because it runs the oracle as the CPU yardstick, which only test infrastructure may do.

    python tests/bench_inverse.py [scale]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # oracle_lib


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    from bifromq_b200 import retain, workload
    w = workload.Workload("C5", scale=scale)
    idx = retain.GpuTopicMatchIndex(0)
    tenants = w.tenants
    t0 = time.perf_counter()
    ids = idx.add_blobs(tenants, w.topics, w.topic_off, w.topic_tenant[:w.n_topics])
    idx.commit()
    build_s = time.perf_counter() - t0
    out = {"config": "C5", "scale": scale, "retained_topics": w.n_topics, "filters": w.n_query_filters, "build_s": round(build_s, 2)}
    for name, lim in (("limit10", np.full(w.n_query_filters, 10, np.int64)), ("unlimited", None)):
        for _ in range(2):
            idx.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant[:w.n_query_filters], lim)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            r = idx.match_blobs(tenants, w.filters, w.filter_off, w.filter_tenant[:w.n_query_filters], lim)
            ts.append(time.perf_counter() - t0)
        out[name] = {"filters_per_s": w.n_query_filters / min(ts), "ids_returned": int(len(r.ids)), "matches_total": int(r.totals.sum()),
                     "ms": round(min(ts) * 1e3, 3), "breakdown_ms": r.timings_ms}
    # oracle on the host cores (bounded sample of the filters, full index)
    import oracle_lib as O
    orc = O.TopicLevelIndex()
    tl = w.topic_list()
    t0 = time.perf_counter()
    for i in range(w.n_topics):
        orc.add(tl[i], int(ids[i]), tenants[w.topic_tenant[i]])
    ns = min(w.n_query_filters, 20000)
    tb, toff = O.blob(tenants)
    counts = np.zeros(ns, np.int64)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.lib.orc_tli_match_batch(orc.h, tb.ctypes.data, toff, np.ascontiguousarray(w.filter_tenant[:ns]).ctypes.data, w.filters.ctypes.data,
                              np.ascontiguousarray(w.filter_off[:ns + 1]), ns, cores, counts, None)
    dt = time.perf_counter() - t0
    out["cpu_oracle"] = {"filters_per_s": ns / dt, "cores": cores, "sample_filters": ns, "matches": int(counts.sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
