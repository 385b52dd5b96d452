package org.apache.bifromq.dist.worker.gpumatch;

import java.nio.ByteBuffer;

/**
 * Thin binding of include/bfq_gpumatch.h (C side: jni/bfq_gpumatch_jni.c, one function per method below).
 * All buffers are DIRECT ByteBuffers (little endian); strings travel as (blob, int64 offsets[n+1]) pairs, exactly the layout the
 * C header documents. Handles are opaque longs. A non-zero C return code surfaces as IllegalStateException carrying
 * bfq_last_error().
 *
 * NOT compiled in this repository (the image has no JDK): this is the file a maintainer adds to bifromq-dist-worker.
 */
final class BfqNative {
    static {
        System.loadLibrary("bfq_gpumatch_jni");   // links libbfq_gpumatch.so
    }

    private BfqNative() {
    }

    // ---- forward index: one per dist-worker KV range (DistWorkerCoProc)
    static native long indexCreate(int deviceOrdinal);                       // bfq_index_create
    static native void indexDestroy(long h);                                 // bfq_index_destroy
    static native void indexReset(long h);                                   // bfq_index_reset       <- DistWorkerCoProc.reset
    static native void indexLoad(long h, ByteBuffer keys, ByteBuffer keyOff, ByteBuffer vals, ByteBuffer valOff, long n);
    static native void indexApply(long h, ByteBuffer addKeys, ByteBuffer addKeyOff, ByteBuffer addVals, ByteBuffer addValOff,
                                  long nAdd, ByteBuffer delKeys, ByteBuffer delKeyOff, long nDel);   // <- mutate()'s Supplier
    static native void indexCommit(long h);                                  // bfq_index_commit
    static native long indexGeneration(long h);                              // bfq_index_generation

    // ---- forward match == ITenantRouteMatcher.matchAll, batched over tenants; thread-safe per handle
    static native long match(long h, ByteBuffer tenants, ByteBuffer tenantOff, int nTenants, ByteBuffer topics,
                             ByteBuffer topicOff, ByteBuffer topicTenant, long nTopics, int[] maxPFanout, int[] maxGFanout);
    static native ByteBuffer resultSpanBegin(long r);    // uint32[nTopics]   views over the result's own pinned memory,
    static native ByteBuffer resultSpanCount(long r);    // uint32[nTopics]   valid until resultFree(r)
    static native ByteBuffer resultRouteCount(long r);   // uint32[nTopics]
    static native ByteBuffer resultRanges(long r);       // {uint32 first, uint32 count}[...]
    static native ByteBuffer resultThrottled(long r);    // {uint32 topic, uint32 rank, uint32 kind}[...]
    static native long[] resultExpand(long r);           // offsets[n+1] ++ surviving ranks (ascending per topic)
    static native byte[][] resultRouteLookup(long r, long rank);   // {key, value} of a rank OF THIS RESULT's snapshot
    static native long resultGeneration(long r);
    static native void resultFree(long r);

    // ---- inverse index: IRetainTopicIndex / TopicIndex
    static native long rindexCreate(int deviceOrdinal);
    static native void rindexDestroy(long h);
    static native long[] rindexAdd(long h, ByteBuffer tenants, ByteBuffer tenantOff, int nTenants, ByteBuffer topics,
                                   ByteBuffer topicOff, ByteBuffer topicTenant, long n);          // -> topic ids
    static native void rindexRemove(long h, byte[] tenant, byte[] topic);
    static native void rindexCommit(long h);
    static native long[] rmatch(long h, ByteBuffer tenants, ByteBuffer tenantOff, int nTenants, ByteBuffer filters,
                                ByteBuffer filterOff, ByteBuffer filterTenant, long n, long[] limit);   // offsets[n+1] ++ ids
    // RetainStoreCoProc.load: feed the raw retain-store KV keys of the range scan (no TopicMessage parsing) -> ids, -1 = not a key
    static native long[] rindexLoadKeys(long h, ByteBuffer keys, ByteBuffer keyOff, long n);
    // RetainStoreCoProc.match: as rmatch, but the matched topics come back as the retain KV keys to reader.get():
    // {long[] offsets[n+1] of filters -> key index, long[] keyOff[nKeys+1], byte[] keyBlob}
    static native Object[] rmatchRetainKeys(long h, ByteBuffer tenants, ByteBuffer tenantOff, int nTenants, ByteBuffer filters,
                                            ByteBuffer filterOff, ByteBuffer filterTenant, long n, long[] limit);

    // ---- dist-server side: TenantRangeLookupCache.lookup for a whole batch -> keepOff[nTopics+1] ++ keep flags (one per candidate)
    static native long[] rangeLookup(int deviceOrdinal, ByteBuffer tenants, ByteBuffer tenantOff, int nTenants, ByteBuffer topics,
                                     ByteBuffer topicOff, ByteBuffer topicTenant, long nTopics, ByteBuffer candOff,
                                     ByteBuffer candFlags, ByteBuffer firstBlob, ByteBuffer firstOff, ByteBuffer lastBlob,
                                     ByteBuffer lastOff);
}
