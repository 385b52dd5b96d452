package org.apache.bifromq.dist.worker.gpumatch;

import java.util.HashMap;
import java.util.HashSet;
import java.util.Map;
import java.util.Set;
import org.apache.bifromq.retain.store.index.IRetainTopicIndex;
import org.apache.bifromq.retain.store.index.RetainedMsgInfo;

/**
 * Drop-in for RetainTopicIndex behind IRetainTopicIndex (bifromq-retain/bifromq-retain-store/.../index/IRetainTopicIndex.java:
 * 27-35; replaced: RetainTopicIndex.java:35-144 over bifromq-util/.../index/TopicLevelTrie.java:190-249). The topics live in a
 * bfq_rindex (BFS-numbered topic trie in HBM, csrc/rmatch_kernels.cu); this class keeps what the native side does not need:
 * timestamp / expiry per topic id.
 *
 * Mutations are staged natively and published by the next match (rindexCommit) — RetainStoreCoProc applies add / remove from
 * the raft-apply thread and matches from the query executor (RetainStoreCoProc.java:167-190, 198-277), so the methods are
 * synchronized like the pcollections trie's compare-and-set loop is lock-free: same visibility, simpler code. For throughput
 * RetainStoreCoProc.match should hand the WHOLE BatchMatchRequest to BfqNative.rmatch / rmatchRetainKeys in one call (all filters
 * of all tenants, per-filter limit): match(tenantId, filter) below is the one-filter form the interface prescribes.
 *
 * NOT compiled in this repository (no JDK in the image). A maintainer who keeps the two modules apart moves BfqNative's
 * rindex natives into a class of bifromq-retain-store; they are declared package-private here for brevity.
 */
public class GpuRetainTopicIndex implements IRetainTopicIndex, AutoCloseable {
    private final long handle;
    private final Map<Long, RetainedMsgInfo> byId = new HashMap<>();          // stable topic id -> info
    private final Map<String, Map<String, Long>> ids = new HashMap<>();       // tenant -> topic -> id (for remove / overwrite)
    private boolean dirty = false;

    public GpuRetainTopicIndex(int deviceOrdinal) {
        handle = BfqNative.rindexCreate(deviceOrdinal);
    }

    @Override
    public synchronized void add(String tenantId, String topic, long timestamp, int expirySeconds) {
        Blobs tenants = Blobs.ofUtf8(tenantId);
        Blobs topics = Blobs.ofUtf8(topic);
        // adding an existing (tenant, topic) returns its id: the info is replaced, like TopicLevelTrie.add on an equal value
        long id = BfqNative.rindexAdd(handle, tenants.blob, tenants.off, 1, topics.blob, topics.off, topics.zeros, 1)[0];
        byId.put(id, new RetainedMsgInfo(tenantId, topic, timestamp, expirySeconds));
        ids.computeIfAbsent(tenantId, t -> new HashMap<>()).put(topic, id);
        dirty = true;
    }

    @Override
    public synchronized void remove(String tenantId, String topic) {
        Map<String, Long> ofTenant = ids.get(tenantId);
        Long id = ofTenant == null ? null : ofTenant.remove(topic);
        if (id == null) {
            return;
        }
        if (ofTenant.isEmpty()) {
            ids.remove(tenantId);
        }
        byId.remove(id);
        BfqNative.rindexRemove(handle, tenantId.getBytes(java.nio.charset.StandardCharsets.UTF_8),
            topic.getBytes(java.nio.charset.StandardCharsets.UTF_8));
        dirty = true;
    }

    @Override
    public synchronized Set<RetainedMsgInfo> match(String tenantId, String topicFilter) {
        publish();
        Blobs tenants = Blobs.ofUtf8(tenantId);
        Blobs filters = Blobs.ofUtf8(topicFilter);
        // no limit here: the interface returns the whole set and RetainStoreCoProc.match stops after `limit` reads (:177-188);
        // the batched entry point takes the limits and truncates on the device
        long[] r = BfqNative.rmatch(handle, tenants.blob, tenants.off, 1, filters.blob, filters.off, filters.zeros, 1, null);
        Set<RetainedMsgInfo> out = new HashSet<>();
        for (int i = 2; i < r.length; i++) {      // r = offsets[0..1] ++ ids
            RetainedMsgInfo info = byId.get(r[i]);
            if (info != null) {
                out.add(info);
            }
        }
        return out;
    }

    @Override
    public synchronized Set<RetainedMsgInfo> findAll() {
        return new HashSet<>(byId.values());
    }

    private void publish() {
        if (dirty) {
            BfqNative.rindexCommit(handle);
            dirty = false;
        }
    }

    @Override
    public synchronized void close() {
        BfqNative.rindexDestroy(handle);
    }
}
