package org.apache.bifromq.dist.worker.gpumatch;

import static org.apache.bifromq.plugin.eventcollector.ThreadLocalEventPool.getLocal;
import static com.google.protobuf.UnsafeByteOperations.unsafeWrap;

import io.micrometer.core.instrument.Timer;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.HashMap;
import java.util.Map;
import java.util.Set;
import org.apache.bifromq.dist.worker.cache.IMatchedRoutes;
import org.apache.bifromq.dist.worker.cache.ITenantRouteMatcher;
import org.apache.bifromq.dist.worker.cache.MatchedRoutes;
import org.apache.bifromq.dist.worker.schema.KVSchemaUtil;
import org.apache.bifromq.dist.worker.schema.cache.GroupMatching;
import org.apache.bifromq.dist.worker.schema.cache.Matching;
import org.apache.bifromq.dist.worker.schema.cache.NormalMatching;
import org.apache.bifromq.plugin.eventcollector.IEventCollector;
import org.apache.bifromq.plugin.eventcollector.distservice.GroupFanoutThrottled;
import org.apache.bifromq.plugin.eventcollector.distservice.PersistentFanoutThrottled;

/**
 * Drop-in for TenantRouteMatcher behind the same seam (ITenantRouteMatcher.java:28-38): an entry for every requested topic,
 * caps applied in KV order, throttle events reported. Created by TenantRouteCacheFactory.create (:67-71) with the range's
 * bfq_index handle instead of a KV reader supplier.
 *
 * Safe to call from the shared topic-matcher ForkJoinPool (DistWorkerCoProcFactory.java:74-85): every bfq_match leases its own
 * workspace and pins the snapshot it ran on; ranks are re-hydrated through resultRouteLookup, i.e. against THAT snapshot, so a
 * commit from the raft-apply thread between the match and the lookups cannot shift them.
 *
 * NOT compiled in this repository (no JDK in the image).
 */
public class GpuTenantRouteMatcher implements ITenantRouteMatcher {
    private final String tenantId;
    private final long index;
    private final IEventCollector eventCollector;
    private final Timer timer;

    public GpuTenantRouteMatcher(String tenantId, long rangeIndexHandle, IEventCollector eventCollector, Timer timer) {
        this.tenantId = tenantId;
        this.index = rangeIndexHandle;
        this.eventCollector = eventCollector;
        this.timer = timer;
    }

    @Override
    public Map<String, IMatchedRoutes> matchAll(Set<String> topics, int maxPersistentFanoutCount, int maxGroupFanoutCount) {
        Timer.Sample sample = Timer.start();
        String[] ts = topics.toArray(new String[0]);
        Blobs tenant = Blobs.ofUtf8(tenantId);
        Blobs b = Blobs.ofUtf8(ts);                                   // direct buffers: blob + int64 offsets; topicTenant = zeros
        long r = BfqNative.match(index, tenant.blob, tenant.off, 1, b.blob, b.off, b.zeros, ts.length,
            new int[] {maxPersistentFanoutCount}, new int[] {maxGroupFanoutCount});
        try {
            long[] csr = BfqNative.resultExpand(r);                   // offsets[n+1] ++ surviving ranks, ascending per topic
            Map<Long, Matching> byRank = new HashMap<>();             // valid for THIS result only: ranks shift with every commit
            Map<String, IMatchedRoutes> out = new HashMap<>();
            for (int i = 0; i < ts.length; i++) {
                MatchedRoutes m = new MatchedRoutes(tenantId, ts[i], eventCollector, Integer.MAX_VALUE, Integer.MAX_VALUE);
                for (long j = csr[i]; j < csr[i + 1]; j++) {
                    Matching matching = byRank.computeIfAbsent(csr[ts.length + 1 + (int) j], rank -> {
                        byte[][] kv = BfqNative.resultRouteLookup(r, rank);
                        return KVSchemaUtil.buildMatchRoute(unsafeWrap(kv[0]), unsafeWrap(kv[1]));
                    });
                    if (matching.type() == Matching.Type.Normal) {
                        m.addNormalMatching((NormalMatching) matching);
                    } else {
                        m.putGroupMatching((GroupMatching) matching);
                    }
                }
                m.adjust(maxPersistentFanoutCount, maxGroupFanoutCount);   // restore the real limits (nothing is over them)
                out.put(ts[i], m);
            }
            ByteBuffer thr = BfqNative.resultThrottled(r).order(ByteOrder.LITTLE_ENDIAN);   // {topic, rank, kind}
            while (thr.hasRemaining()) {
                int topic = thr.getInt();
                long rank = Integer.toUnsignedLong(thr.getInt());
                int kind = thr.getInt();
                byte[][] kv = BfqNative.resultRouteLookup(r, rank);
                String mqttTopicFilter = KVSchemaUtil.buildMatchRoute(unsafeWrap(kv[0]), unsafeWrap(kv[1])).mqttTopicFilter();
                if (kind == 1) {
                    eventCollector.report(getLocal(PersistentFanoutThrottled.class).tenantId(tenantId).topic(ts[topic])
                        .mqttTopicFilter(mqttTopicFilter).maxCount(maxPersistentFanoutCount));
                } else {
                    eventCollector.report(getLocal(GroupFanoutThrottled.class).tenantId(tenantId).topic(ts[topic])
                        .mqttTopicFilter(mqttTopicFilter).maxCount(maxGroupFanoutCount));
                }
            }
            sample.stop(timer);                                       // the reference's "dist.match.internal" timer
            return out;
        } finally {
            BfqNative.resultFree(r);
        }
    }
}
