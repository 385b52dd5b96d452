#!/usr/bin/env python
"""bench.py — publish-topic matching throughput of the CUDA matcher (and the CPU reference arm).

One "step" = one pass of the hot path over one batch of synthetic publish topics:
    value  = topics/s with the batch resident in HBM (bfq_match_device; kernels + counter read-back)
    e2e    = topics/s through the host-buffer C-ABI call bfq_match (pinned host -> H2D -> kernels -> D2H result)
Workload = BASELINE.json config C4 by default (10M filters over 1000 tenants, Zipf-skewed fan-out, 1M-topic batch):
the metric is quoted "@10M filters" and it fits one B200. Under torchrun every rank owns its own tenants
(tenant sharding, no data-path collective; weak scaling: each rank hosts a full-size shard) unless --scaling strong.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node 8 ... bench.py --gpus 8
    python bench.py --impl reference        # the reference algorithm restated in C++ (oracle/), on host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "publish-topics matched/sec @10M filters"
UNIT = "topics/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; invalid as a bench number)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="topics in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (the sampler is started early so that its
    first samples exist before the region begins; rows are then filtered by timestamp)."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        """gpu_index: one index or a comma-separated list (rank 0 samples every GPU of the job from ONE nvidia-smi process:
        eight pollers at 20 ms contend for the driver and slow the ranks' launches)"""
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 5.0:   # wait for the first sample
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        inside = [r for ts, r in self.rows if self.t_begin is not None and self.t_begin - 0.02 <= ts <= (self.t_end or ts) + 0.04]
        rows = inside or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside)}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def pin_to_gpu_numa_node(dev_index):
    """Run this rank on the CPU socket its GPU hangs off (and first-touch its pinned buffers there): on a two-socket box a
    process that lands on the far socket sees a third less PCIe bandwidth, which is what the end-to-end number measures.
    Returns {"node", "cpus", "previous"} or None when the topology is not exposed; never raises."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(dev_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        previous = os.sched_getaffinity(0)
        target = cpus & previous
        if not target or target == previous:
            return None
        os.sched_setaffinity(0, target)
        return {"node": node, "cpus": len(target), "previous": previous}
    except Exception:
        return None


def make_roofline(sample_topic_bytes, st, ns, n_topics_per_launch, kernel_ms):
    """`roofline` object of the JSON line. SURVEY.md §8(d): algorithmic bytes per topic
    B = len(topic) + 4 + 32 V + 8 P + 8 ranges (range-encoded output) + 4, with V / P / ranges counted by the oracle on the
    cpu_baseline sample (`st`, over `ns` topics); achieved = B x topics per launch / the tier-0 kernel's duration."""
    per_topic = (sample_topic_bytes + 4 * ns + 32 * st["V"] + 8 * st["P"] + 8 * st["ranges"] + 4 * ns) / ns
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = per_topic * n_topics_per_launch / (kernel_ms / 1000.0) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_kernel_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
            "frac_of_nominal_8000": achieved / 8000.0,
            "kernel": "match_topics_lane_kernel (tier 0, one lane per topic)", "kernel_ms": kernel_ms, "alg_bytes_per_topic": per_topic,
            "alg_counters_per_topic": {"V": st["V"] / ns, "P": st["P"] / ns, "ranges": st["ranges"] / ns, "R": st["R"] / ns},
            "note": "algorithmic bytes per topic measured by the oracle on the cpu_baseline sample"}
    if traffic:
        # DRAM bytes the kernel actually moved (ncu capture of the same command) over the live time
        roof["dram_gbs_from_ncu_traffic"] = traffic / (kernel_ms / 1000.0) / 1e9
        roof["traffic_over_algorithmic"] = traffic / (per_topic * n_topics_per_launch)
    return roof


def make_workload(args, rank, world):
    from bifromq_b200.workload import Workload
    if world > 1 and args.scaling == "strong":
        return Workload(args.config, scale=args.scale, shard_index=rank, shard_count=world)
    # weak scaling: every rank hosts a full-size shard with its own tenant namespace
    prefix = "" if world == 1 else "g%d-" % rank
    return Workload(args.config, seed=Workload.SEED + rank, scale=args.scale, tenant_prefix=prefix)


def cpu_sample_indices(w, want):
    """bounded sample of the batch: every topic of every 8th tenant (keeps the largest tenant), capped at `want`"""
    tt = np.asarray(w.topic_tenant[:w.n_topics])
    keep = np.nonzero(tt % 8 == 0)[0] if w.n_tenants >= 8 else np.arange(w.n_topics)
    if len(keep) > want:
        keep = keep[np.linspace(0, len(keep) - 1, want).astype(np.int64)]
    return keep


def oracle_for_sample(w, idx):
    """load only the sampled tenants' routes into the oracle (tenants are independent key ranges)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    tt = np.asarray(w.topic_tenant[:w.n_topics])[idx]
    tenants = w.tenants
    used = sorted(set(tt.tolist()))
    kv = O.KV()
    keys_all = w.keys
    # tenant key ranges: binary search on the tenant begin keys
    kb = keys_all.tobytes() if w.n_routes < 3_000_000 else None
    for t in used:
        begin = O.tenant_begin_key(tenants[t])
        end = O.upper_bound(begin)
        lo = _lower_bound(w, begin, kb)
        hi = _lower_bound(w, end, kb)
        if hi > lo:
            ko = np.ascontiguousarray(w.key_off[lo:hi + 1])
            vo = np.ascontiguousarray(w.val_off[lo:hi + 1])
            O.lib.orc_kv_load(kv.h, w.keys.ctypes.data, ko, w.vals.ctypes.data, vo, hi - lo)
    kv.freeze()
    remap = {t: i for i, t in enumerate(used)}
    sub_tenants = [tenants[t] for t in used]
    sub_tt = np.array([remap[t] for t in tt.tolist()], np.int32)
    topics = [w.topic(int(i)) for i in idx]
    return O, kv, sub_tenants, topics, sub_tt


def _lower_bound(w, key, kb):
    lo, hi = 0, w.n_routes
    mv = memoryview(w.keys)
    while lo < hi:
        mid = (lo + hi) // 2
        k = bytes(mv[w.key_off[mid]:w.key_off[mid + 1]])
        if k < key:
            lo = mid + 1
        else:
            hi = mid
    return lo


def run_cpu_baseline(w, args, mode_name):
    """times the oracle on the host cores over a bounded sample; returns the cpu_baseline dict and the per-topic
    algorithmic-byte figures measured on the same sample"""
    cores = os.cpu_count() or 1
    want = args.cpu_sample or 200000
    idx = cpu_sample_indices(w, want)
    O, kv, tenants, topics, tt = oracle_for_sample(w, idx)
    tb, toff = O.blob(tenants)
    pb, poff = O.blob(topics)
    mode = O.MODE_TRIE if mode_name == "trie" else O.MODE_REFERENCE
    singleton = mode_name != "trie"
    # warm (also builds the oracle's trie outside the timed region)
    kv.match_blobs(tb, toff, pb, poff, tt, min(len(topics), 256), 2 ** 31 - 1, 100, mode, singleton, cores)
    passes, dt = (5 if mode_name == "trie" else 1), 0.0
    for _ in range(passes):
        out = kv.match_blobs(tb, toff, pb, poff, tt, len(topics), 2 ** 31 - 1, 100, mode, singleton, cores)
        dt += kv.last_match_seconds   # the C++ matcher call alone (result marshalling to numpy excluded)
    dt /= passes
    n = len(topics)
    stats = out.stats
    res = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d topics (all topics of every 8th tenant, evenly thinned) against those tenants' %d routes; %s; %.2f s wall per pass = %.0f core-seconds"
                     % (n, len(kv), "oracle filter-trie walk, std::thread x %d, mean of 5 passes" % cores if mode_name == "trie" else
                        "literal TenantRouteMatcher.matchAll restatement, one call per topic (production shape), std::thread x %d" % cores,
                        dt, dt * cores)}
    return res, stats, n, float(np.diff(poff).sum())


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world

    if args.impl == "reference":
        # the reference's own algorithm (restated in C++, oracle/) on the host cores; rank 0 only
        if rank != 0:
            return
        w = make_workload(args, 0, 1)
        samples, per = [], None
        base, _, n, _ = run_cpu_baseline(w, args, "reference")
        # the contract's K steps: each step is the same bounded sample; W warm-ups are untimed
        vals = [base["value"]]
        for _ in range(max(0, min(args.steps, 3) - 1)):
            b2, _, _, _ = run_cpu_baseline(w, args, "reference")
            vals.append(b2["value"])
        v = float(np.mean(vals))
        base["value"] = v
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 1,
                "ms_per_step": 1000.0 * n / v, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "u8/u32 (byte and integer work)", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload_name(args, w), "note": "C++ restatement of the Java reference, not the JVM (no JDK in the image)"},
                "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    import bifromq_b200
    bifromq_b200.load_library()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)

    t_gen = time.perf_counter()
    w = make_workload(args, rank, world)
    t_gen = time.perf_counter() - t_gen
    t_build = time.perf_counter()
    idx = bifromq_b200.GpuRouteIndex(local)
    idx.load(w.keys, w.key_off, w.vals, w.val_off)
    idx.commit()
    t_build = time.perf_counter() - t_build
    stats = idx.stats()
    tenants = idx.tenant_blob(w.tenants)   # marshalled once: the same tenant list serves every batch
    n = w.n_topics
    blob_bytes = int(w.topic_off[-1])

    # ---- device-resident batch (value) and pinned host batch (e2e)
    h_topics = torch.from_numpy(np.ascontiguousarray(w.topics[:max(blob_bytes, 1)])).pin_memory()
    h_off = torch.from_numpy(np.ascontiguousarray(w.topic_off)).pin_memory()
    h_tt = torch.from_numpy(np.ascontiguousarray(w.topic_tenant[:max(n, 1)])).pin_memory()
    d_topics, d_off, d_tt = h_topics.to(dev), h_off.to(dev), h_tt.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def step_device():
        return idx.match_device(tenants, d_topics.data_ptr(), d_off.data_ptr(), d_tt.data_ptr(), n, stream=stream.cuda_stream)

    def step_e2e():
        r = idx.match(tenants, h_topics.numpy(), h_off.numpy(), h_tt.numpy())
        d2h = 12 * n + 8 * len(r.ranges) + 12 * len(r.throttled)
        tm = r.timings_ms
        r.close()
        return d2h, tm

    sampler = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler.begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, launches, n_ranges, n_overflow = [], 0, 0, 0
    torch.cuda.synchronize(dev)
    for i in range(args.steps):
        flush.zero_()                     # L2 flush between timed iterations (untimed)
        ev[i][0].record(stream)
        out = step_device()
        ev[i][1].record(stream)
        kernel_ms.append(idx.last_kernel_ms())
        launches += out.n_launches
        n_ranges, n_overflow = out.n_ranges, out.n_overflow_topics
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    # ---- the one exchange step of the sharded path: all ranks gather the per-topic fan-out counts of the
    # whole job (what BatchDistReply carries); timed on its own, the matching itself needs no collective
    exchange_ms, imbalance, exchange_error = None, None, None
    if world > 1:
        from bifromq_b200 import dist as D
        try:   # every rank runs the same code on the same shapes, so a failure here is the same failure on every rank
            fan = D.device_view(out.d_route_count, n, "<i4", dev)
            for _ in range(3):
                D.gather_fanout(fan)
            torch.cuda.synchronize(dev)
            dist.barrier()
            xe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            xe[0].record(stream)
            for _ in range(args.steps):
                gathered = D.gather_fanout(fan)
            xe[1].record(stream)
            torch.cuda.synchronize(dev)
            exchange_ms = xe[0].elapsed_time(xe[1]) / args.steps
            assert gathered.numel() == n * world
        except Exception as ex:   # the exchange is reported beside the metric, it must not take the metric down
            exchange_ms, exchange_error = None, "%s: %s" % (type(ex).__name__, ex)
        ex_max, _ = D.aggregate(exchange_ms if exchange_ms is not None else -1.0, 0, dev)
        t_max, _ = D.aggregate(total_ms, 0, dev)
        t_sum = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_sum)
        exchange_ms, imbalance = (ex_max if exchange_ms is not None else None), t_max / (t_sum.item() / world)
    n_routes = int(torch.from_numpy(np.zeros(1)).sum()) if n == 0 else None
    # ---- e2e through the host-buffer call
    for _ in range(2):
        step_e2e()
    e2e_t, d2h_bytes, e2e_tm = [], 0, {}
    if world > 1:
        dist.barrier()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        d2h_bytes, e2e_tm = step_e2e()
        e2e_t.append(time.perf_counter() - t0)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    if numa:   # the CPU baseline below uses every host core
        try:
            os.sched_setaffinity(0, numa["previous"])
        except Exception:
            pass
    e2e_total = float(sum(e2e_t))
    h2d_bytes = blob_bytes + 8 * (n + 1) + 4 * n

    # ---- whole-job numbers: MAX over ranks of the time, SUM over ranks of the topics
    from bifromq_b200 import dist as D
    total_ms_max, topics_all = D.aggregate(total_ms, n, dev)
    e2e_ms_max, _ = D.aggregate(e2e_total * 1000.0, n, dev)
    value = topics_all * args.steps / (total_ms_max / 1000.0)
    e2e_value = topics_all * args.steps / (e2e_ms_max / 1000.0)

    if rank == 0:
        # matched routes of one batch (for the fan-out routes/s figure)
        res = idx.match(tenants, h_topics.numpy(), h_off.numpy(), h_tt.numpy())
        routes_per_batch = int(res.route_count.astype(np.int64).sum())
        res.close()
        stats = idx.stats()
        cpu_base, roof = None, None
        k_ms = float(np.mean(kernel_ms))
        if not args.no_cpu_baseline:
            cpu_base, st, ns, sample_topic_bytes = run_cpu_baseline(w, args, "trie")
            roof = make_roofline(sample_topic_bytes, st, ns, n, k_ms)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "u8/u32 (byte and integer work)", "data": "synthetic",
                "config": {"workload": workload_name(args, w), "routes_per_gpu": w.n_routes, "filters_per_gpu": w.n_filters,
                           "tenants_per_gpu": w.n_tenants, "topics_per_step_per_gpu": n, "parallelism": "tenant-sharded x%d" % world,
                           "l2": "flushed between timed steps (256 MiB memset, untimed); index %.2f GB >> L2" % (stats["device_bytes"] / 1e9),
                           "caps": "MaxPersistentFanout=INT_MAX, MaxGroupFanout=INT_MAX",
                           "order": "tier 0 picks the topics in locality order (order_keys_kernel + cub radix sort, inside the timed region)",
                           "host": ("rank pinned to NUMA node %d of its GPU (%d cpus) for the GPU legs" % (numa["node"], numa["cpus"])) if numa
                                   else "no NUMA pinning (topology not exposed or single node)",
                           "gen_s": round(t_gen, 1), "build_s": round(t_build, 1)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "last_step_breakdown_ms": {k: round(v, 3) for k, v in e2e_tm.items()}},
                "gpu_launches": launches, "gpu_launches_note": "own kernels only: order_keys + tier 0 + tier 1 per step (cub's sort passes not counted)", "routes_per_s": routes_per_batch * world * args.steps / (total_ms_max / 1000.0),
                "ranges_per_step": n_ranges, "tier2_topics_per_step": n_overflow, "index": stats, "clocks": clocks}
        if exchange_error is not None:
            line["exchange"] = {"error": exchange_error}
            line["load_imbalance_max_over_mean"] = imbalance
        if exchange_ms is not None:
            line["exchange"] = {"what": "all-gather of per-topic fan-out counts (int32) over NCCL, all ranks end with the whole job's",
                                "ms_per_step": exchange_ms, "bytes_per_rank": 4 * n,
                                "value_with_exchange": topics_all * args.steps / ((total_ms_max + exchange_ms * args.steps) / 1000.0)}
            line["load_imbalance_max_over_mean"] = imbalance
        if roof:
            line["roofline"] = roof
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def workload_name(args, w):
    names = {"C1": "C1: 1 tenant, 10k exact filters, 1k topics", "C2": "C2: 1 tenant, 1M filters (50% '+'), 100k-topic batch",
             "C3": "C3: 1000 tenants x 10k filters mixed +/#, 1M-topic batch",
             "C4": "C4: 10M filters over 1000 tenants (Zipf sizes, Zipf fan-out and topic popularity), 1M-topic batch"}
    s = names[args.config]
    if args.scale != 1.0:
        s += " [scale %.4g — NOT a valid bench size]" % args.scale
    return s


if __name__ == "__main__":
    main()
