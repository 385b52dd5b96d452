/* bfq_gpumatch_jni.c — the JNI shim a bifromq maintainer adds: org.apache.bifromq.dist.worker.gpumatch.BfqNative
 * (jni/java/.../BfqNative.java) over include/bfq_gpumatch.h. One function per native method: unwrap the arguments, call the
 * C-ABI, turn a non-zero code into an IllegalStateException carrying bfq_last_error() — which DistWorkerCoProc.query already
 * maps to an exceptionally completed future (bifromq-dist/bifromq-dist-worker/src/main/java/org/apache/bifromq/dist/worker/
 * DistWorkerCoProc.java:142-152). No state lives here: handles are jlong-wrapped pointers, buffers are DIRECT ByteBuffers
 * (little endian) in the (blob, int64 offsets[n+1]) layout the header documents.
 *
 * Build (maintainer):  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include bfq_gpumatch_jni.c \
 *                          -L../bifromq_b200 -lbfq_gpumatch -o libbfq_gpumatch_jni.so
 * Here (no JDK in the image): type-checked against jni_stub.h with -DBFQ_JNI_STUB by __graft_entry__.build().
 */
#ifdef BFQ_JNI_STUB
#include "jni_stub.h"
#else
#include <jni.h>
#endif
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bfq_gpumatch.h"

#define JFN(name) JNIEXPORT JNICALL Java_org_apache_bifromq_dist_worker_gpumatch_BfqNative_##name
#define ADDR(buf) ((buf) ? (*env)->GetDirectBufferAddress(env, (buf)) : NULL)
#define IDX(h) ((bfq_index*) (intptr_t) (h))
#define RES(r) ((bfq_result*) (intptr_t) (r))
#define RIDX(h) ((bfq_rindex*) (intptr_t) (h))

static void throw_bfq(JNIEnv* env, int32_t rc) {
    char msg[768];
    snprintf(msg, sizeof msg, "bfq error %d: %s", (int) rc, bfq_last_error());
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalStateException"), msg);
}
static jobject view(JNIEnv* env, const void* p, jlong bytes) {
    return (*env)->NewDirectByteBuffer(env, (void*) p, bytes);   /* read-only by contract: the result owns the memory */
}

/* ---------------------------------------------------------------- forward index life cycle */
jlong JFN(indexCreate)(JNIEnv* env, jclass cls, jint device) {
    (void) cls;
    bfq_index* h = NULL;
    int32_t rc = bfq_index_create(device, &h);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return 0; }
    return (jlong) (intptr_t) h;
}
void JFN(indexDestroy)(JNIEnv* env, jclass cls, jlong h) { (void) env; (void) cls; bfq_index_destroy(IDX(h)); }
void JFN(indexReset)(JNIEnv* env, jclass cls, jlong h) {
    (void) cls;
    int32_t rc = bfq_index_reset(IDX(h));
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
void JFN(indexLoad)(JNIEnv* env, jclass cls, jlong h, jobject keys, jobject keyOff, jobject vals, jobject valOff, jlong n) {
    (void) cls;
    int32_t rc = bfq_index_load(IDX(h), ADDR(keys), ADDR(keyOff), ADDR(vals), ADDR(valOff), n);
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
void JFN(indexApply)(JNIEnv* env, jclass cls, jlong h, jobject addKeys, jobject addKeyOff, jobject addVals, jobject addValOff,
                     jlong nAdd, jobject delKeys, jobject delKeyOff, jlong nDel) {
    (void) cls;
    int32_t rc = bfq_index_apply(IDX(h), ADDR(addKeys), ADDR(addKeyOff), ADDR(addVals), ADDR(addValOff), nAdd, ADDR(delKeys),
                                 ADDR(delKeyOff), nDel);
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
void JFN(indexCommit)(JNIEnv* env, jclass cls, jlong h) {
    (void) cls;
    int32_t rc = bfq_index_commit(IDX(h));
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
jlong JFN(indexGeneration)(JNIEnv* env, jclass cls, jlong h) {
    (void) cls;
    uint64_t g = 0;
    int32_t rc = bfq_index_generation(IDX(h), &g);
    if (rc != BFQ_OK) throw_bfq(env, rc);
    return (jlong) g;
}

/* ---------------------------------------------------------------- forward match */
jlong JFN(match)(JNIEnv* env, jclass cls, jlong h, jobject tenants, jobject tenantOff, jint nTenants, jobject topics,
                 jobject topicOff, jobject topicTenant, jlong nTopics, jintArray maxP, jintArray maxG) {
    (void) cls;
    jint* p = maxP ? (*env)->GetIntArrayElements(env, maxP, NULL) : NULL;
    jint* g = maxG ? (*env)->GetIntArrayElements(env, maxG, NULL) : NULL;
    bfq_result* r = NULL;
    int32_t rc = bfq_match(IDX(h), ADDR(tenants), ADDR(tenantOff), nTenants, ADDR(topics), ADDR(topicOff), ADDR(topicTenant), nTopics,
                           (const int32_t*) p, (const int32_t*) g, &r);
    if (p) (*env)->ReleaseIntArrayElements(env, maxP, p, JNI_ABORT);
    if (g) (*env)->ReleaseIntArrayElements(env, maxG, g, JNI_ABORT);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return 0; }
    return (jlong) (intptr_t) r;
}
/* views over the result's own pinned memory: valid until resultFree, whatever else runs on the handle */
jobject JFN(resultSpanBegin)(JNIEnv* env, jclass cls, jlong r) { (void) cls; return view(env, bfq_result_span_begin(RES(r)), 4 * bfq_result_num_topics(RES(r))); }
jobject JFN(resultSpanCount)(JNIEnv* env, jclass cls, jlong r) { (void) cls; return view(env, bfq_result_span_count(RES(r)), 4 * bfq_result_num_topics(RES(r))); }
jobject JFN(resultRouteCount)(JNIEnv* env, jclass cls, jlong r) { (void) cls; return view(env, bfq_result_route_count(RES(r)), 4 * bfq_result_num_topics(RES(r))); }
jobject JFN(resultRanges)(JNIEnv* env, jclass cls, jlong r) {
    (void) cls;
    int64_t n = 0;
    const bfq_range* p = bfq_result_ranges(RES(r), &n);
    return view(env, p, (jlong) sizeof(bfq_range) * n);
}
jobject JFN(resultThrottled)(JNIEnv* env, jclass cls, jlong r) {
    (void) cls;
    int64_t n = 0;
    const bfq_throttled* p = bfq_result_throttled(RES(r), &n);
    return view(env, p, (jlong) sizeof(bfq_throttled) * n);
}
/* offsets[n + 1] followed by the surviving ranks, ascending per topic */
jlongArray JFN(resultExpand)(JNIEnv* env, jclass cls, jlong r) {
    (void) cls;
    const int64_t n = bfq_result_num_topics(RES(r));
    int64_t* off = (int64_t*) malloc((size_t) (n + 1) * sizeof(int64_t));
    if (!off) { throw_bfq(env, BFQ_E_NOMEM); return NULL; }
    const int64_t total = bfq_result_expand(RES(r), off, NULL, 0);
    int64_t* all = total >= 0 ? (int64_t*) realloc(off, (size_t) (n + 1 + total) * sizeof(int64_t)) : NULL;
    if (!all) { free(off); throw_bfq(env, total < 0 ? (int32_t) total : BFQ_E_NOMEM); return NULL; }
    bfq_result_expand(RES(r), all, all + n + 1, total);
    jlongArray out = (*env)->NewLongArray(env, (jsize) (n + 1 + total));
    if (out) (*env)->SetLongArrayRegion(env, out, 0, (jsize) (n + 1 + total), (const jlong*) all);
    free(all);
    return out;
}
/* {key, value} of a rank OF THIS RESULT (resolved against the snapshot the match ran on) */
jobjectArray JFN(resultRouteLookup)(JNIEnv* env, jclass cls, jlong r, jlong rank) {
    (void) cls;
    int64_t kl = 0, vl = 0;
    int32_t rc = bfq_result_route_lookup(RES(r), rank, NULL, 0, &kl, NULL, 0, &vl);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return NULL; }
    uint8_t* kb = (uint8_t*) malloc((size_t) (kl + vl + 1));
    if (!kb) { throw_bfq(env, BFQ_E_NOMEM); return NULL; }
    rc = bfq_result_route_lookup(RES(r), rank, kb, kl, &kl, kb + kl, vl, &vl);
    if (rc != BFQ_OK) { free(kb); throw_bfq(env, rc); return NULL; }
    jbyteArray k = (*env)->NewByteArray(env, (jsize) kl), v = (*env)->NewByteArray(env, (jsize) vl);
    (*env)->SetByteArrayRegion(env, k, 0, (jsize) kl, (const jbyte*) kb);
    (*env)->SetByteArrayRegion(env, v, 0, (jsize) vl, (const jbyte*) (kb + kl));
    free(kb);
    jobjectArray out = (*env)->NewObjectArray(env, 2, (*env)->FindClass(env, "[B"), NULL);
    (*env)->SetObjectArrayElement(env, out, 0, k);
    (*env)->SetObjectArrayElement(env, out, 1, v);
    return out;
}
jlong JFN(resultGeneration)(JNIEnv* env, jclass cls, jlong r) { (void) env; (void) cls; return (jlong) bfq_result_generation(RES(r)); }
void JFN(resultFree)(JNIEnv* env, jclass cls, jlong r) { (void) env; (void) cls; bfq_result_free(RES(r)); }

/* ---------------------------------------------------------------- inverse index (retain store / TopicIndex) */
jlong JFN(rindexCreate)(JNIEnv* env, jclass cls, jint device) {
    (void) cls;
    bfq_rindex* h = NULL;
    int32_t rc = bfq_rindex_create(device, &h);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return 0; }
    return (jlong) (intptr_t) h;
}
void JFN(rindexDestroy)(JNIEnv* env, jclass cls, jlong h) { (void) env; (void) cls; bfq_rindex_destroy(RIDX(h)); }
jlongArray JFN(rindexAdd)(JNIEnv* env, jclass cls, jlong h, jobject tenants, jobject tenantOff, jint nTenants, jobject topics,
                          jobject topicOff, jobject topicTenant, jlong n) {
    (void) cls;
    int64_t* ids = (int64_t*) malloc((size_t) (n > 0 ? n : 1) * sizeof(int64_t));
    if (!ids) { throw_bfq(env, BFQ_E_NOMEM); return NULL; }
    int32_t rc = bfq_rindex_add(RIDX(h), ADDR(tenants), ADDR(tenantOff), nTenants, ADDR(topics), ADDR(topicOff), ADDR(topicTenant), n, ids);
    if (rc != BFQ_OK) { free(ids); throw_bfq(env, rc); return NULL; }
    jlongArray out = (*env)->NewLongArray(env, (jsize) n);
    if (out) (*env)->SetLongArrayRegion(env, out, 0, (jsize) n, (const jlong*) ids);
    free(ids);
    return out;
}
void JFN(rindexRemove)(JNIEnv* env, jclass cls, jlong h, jbyteArray tenant, jbyteArray topic) {
    (void) cls;
    jbyte* t = (*env)->GetByteArrayElements(env, tenant, NULL);
    jbyte* p = (*env)->GetByteArrayElements(env, topic, NULL);
    int32_t rc = bfq_rindex_remove(RIDX(h), (const uint8_t*) t, (*env)->GetArrayLength(env, tenant), (const uint8_t*) p, (*env)->GetArrayLength(env, topic));
    (*env)->ReleaseByteArrayElements(env, tenant, t, JNI_ABORT);
    (*env)->ReleaseByteArrayElements(env, topic, p, JNI_ABORT);
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
void JFN(rindexCommit)(JNIEnv* env, jclass cls, jlong h) {
    (void) cls;
    int32_t rc = bfq_rindex_commit(RIDX(h));
    if (rc != BFQ_OK) throw_bfq(env, rc);
}
/* offsets[n + 1] followed by the matched topic ids */
jlongArray JFN(rmatch)(JNIEnv* env, jclass cls, jlong h, jobject tenants, jobject tenantOff, jint nTenants, jobject filters,
                       jobject filterOff, jobject filterTenant, jlong n, jlongArray limit) {
    (void) cls;
    jlong* lim = limit ? (*env)->GetLongArrayElements(env, limit, NULL) : NULL;
    bfq_rresult* r = NULL;
    int32_t rc = bfq_rmatch(RIDX(h), ADDR(tenants), ADDR(tenantOff), nTenants, ADDR(filters), ADDR(filterOff), ADDR(filterTenant), n,
                            (const int64_t*) lim, &r);
    if (lim) (*env)->ReleaseLongArrayElements(env, limit, lim, JNI_ABORT);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return NULL; }
    int64_t n_ids = 0;
    const int64_t* ids = bfq_rresult_ids(r, &n_ids);
    jlongArray out = (*env)->NewLongArray(env, (jsize) (n + 1 + n_ids));
    if (out) {
        (*env)->SetLongArrayRegion(env, out, 0, (jsize) (n + 1), (const jlong*) bfq_rresult_offsets(r));
        if (n_ids) (*env)->SetLongArrayRegion(env, out, (jsize) (n + 1), (jsize) n_ids, (const jlong*) ids);
    }
    bfq_rresult_free(r);
    return out;
}
/* RetainStoreCoProc.load (bifromq-retain/bifromq-retain-store/src/main/java/org/apache/bifromq/retain/store/RetainStoreCoProc.java:
 * 279-296): the raw KV keys of the range scan -> topic ids (-1 = not a retain key) */
jlongArray JFN(rindexLoadKeys)(JNIEnv* env, jclass cls, jlong h, jobject keys, jobject keyOff, jlong n) {
    (void) cls;
    int64_t* ids = (int64_t*) malloc((size_t) (n > 0 ? n : 1) * sizeof(int64_t));
    if (!ids) { throw_bfq(env, BFQ_E_NOMEM); return NULL; }
    int32_t rc = bfq_rindex_load_keys(RIDX(h), ADDR(keys), ADDR(keyOff), n, ids);
    if (rc != BFQ_OK) { free(ids); throw_bfq(env, rc); return NULL; }
    jlongArray out = (*env)->NewLongArray(env, (jsize) n);
    if (out) (*env)->SetLongArrayRegion(env, out, 0, (jsize) n, (const jlong*) ids);
    free(ids);
    return out;
}
/* {offsets[n + 1], keyOff[nKeys + 1], keyBlob}: the matched topics as the retain KV keys RetainStoreCoProc.match reads next (:177-188) */
jobjectArray JFN(rmatchRetainKeys)(JNIEnv* env, jclass cls, jlong h, jobject tenants, jobject tenantOff, jint nTenants, jobject filters,
                                   jobject filterOff, jobject filterTenant, jlong n, jlongArray limit) {
    (void) cls;
    jlong* lim = limit ? (*env)->GetLongArrayElements(env, limit, NULL) : NULL;
    bfq_rresult* r = NULL;
    int32_t rc = bfq_rmatch(RIDX(h), ADDR(tenants), ADDR(tenantOff), nTenants, ADDR(filters), ADDR(filterOff), ADDR(filterTenant), n,
                            (const int64_t*) lim, &r);
    if (lim) (*env)->ReleaseLongArrayElements(env, limit, lim, JNI_ABORT);
    if (rc != BFQ_OK) { throw_bfq(env, rc); return NULL; }
    int64_t n_ids = 0;
    (void) bfq_rresult_ids(r, &n_ids);
    const int64_t blob_len = bfq_rresult_retain_keys(RIDX(h), r, NULL, 0, NULL);
    uint8_t* blob = blob_len >= 0 ? (uint8_t*) malloc((size_t) blob_len + 1) : NULL;
    int64_t* key_off = blob ? (int64_t*) malloc((size_t) (n_ids + 1) * sizeof(int64_t)) : NULL;
    if (!blob || !key_off) {
        free(blob); free(key_off); bfq_rresult_free(r);
        throw_bfq(env, blob_len < 0 ? (int32_t) blob_len : BFQ_E_NOMEM);
        return NULL;
    }
    const int64_t got = bfq_rresult_retain_keys(RIDX(h), r, blob, blob_len, key_off);
    jobjectArray out = NULL;
    if (got == blob_len) {
        jlongArray offs = (*env)->NewLongArray(env, (jsize) (n + 1));
        jlongArray koff = (*env)->NewLongArray(env, (jsize) (n_ids + 1));
        jbyteArray kb = (*env)->NewByteArray(env, (jsize) blob_len);
        (*env)->SetLongArrayRegion(env, offs, 0, (jsize) (n + 1), (const jlong*) bfq_rresult_offsets(r));
        (*env)->SetLongArrayRegion(env, koff, 0, (jsize) (n_ids + 1), (const jlong*) key_off);
        (*env)->SetByteArrayRegion(env, kb, 0, (jsize) blob_len, (const jbyte*) blob);
        out = (*env)->NewObjectArray(env, 3, (*env)->FindClass(env, "java/lang/Object"), NULL);
        (*env)->SetObjectArrayElement(env, out, 0, offs);
        (*env)->SetObjectArrayElement(env, out, 1, koff);
        (*env)->SetObjectArrayElement(env, out, 2, kb);
    } else {
        throw_bfq(env, got < 0 ? (int32_t) got : BFQ_E_STATE);
    }
    free(blob); free(key_off); bfq_rresult_free(r);
    return out;
}

/* ---------------------------------------------------------------- dist-server range pruning (TenantRangeLookupCache.lookup,
 * bifromq-dist/bifromq-dist-server/src/main/java/org/apache/bifromq/dist/server/scheduler/TenantRangeLookupCache.java:70-106)
 * keepOff[nTopics + 1] followed by one 0/1 flag per (topic, candidate range of its tenant) */
jlongArray JFN(rangeLookup)(JNIEnv* env, jclass cls, jint device, jobject tenants, jobject tenantOff, jint nTenants, jobject topics,
                            jobject topicOff, jobject topicTenant, jlong nTopics, jobject candOff, jobject candFlags,
                            jobject firstBlob, jobject firstOff, jobject lastBlob, jobject lastOff) {
    (void) cls;
    const int64_t* coff = (const int64_t*) ADDR(candOff);
    const int32_t* tt = (const int32_t*) ADDR(topicTenant);
    if (!coff || (nTopics > 0 && !tt)) { throw_bfq(env, BFQ_E_INVALID); return NULL; }
    int64_t total = 0;   /* rows are sized by the caller's own candidate table */
    for (jlong i = 0; i < nTopics; i++) {
        if (tt[i] < 0 || tt[i] >= nTenants) { throw_bfq(env, BFQ_E_INVALID); return NULL; }
        total += coff[tt[i] + 1] - coff[tt[i]];
    }
    int64_t* all = (int64_t*) malloc((size_t) (nTopics + 1 + total) * sizeof(int64_t));   /* keepOff ++ widened flags */
    uint8_t* keep = (uint8_t*) malloc((size_t) total + 1);
    if (!all || !keep) { free(all); free(keep); throw_bfq(env, BFQ_E_NOMEM); return NULL; }
    int32_t rc = bfq_range_lookup(device, ADDR(tenants), ADDR(tenantOff), nTenants, ADDR(topics), ADDR(topicOff), tt, nTopics, coff,
                                  ADDR(candFlags), ADDR(firstBlob), ADDR(firstOff), ADDR(lastBlob), ADDR(lastOff), all, keep);
    jlongArray out = NULL;
    if (rc != BFQ_OK) {
        throw_bfq(env, rc);
    } else if ((out = (*env)->NewLongArray(env, (jsize) (nTopics + 1 + total))) != NULL) {
        for (int64_t i = 0; i < total; i++) all[nTopics + 1 + i] = keep[i];
        (*env)->SetLongArrayRegion(env, out, 0, (jsize) (nTopics + 1 + total), (const jlong*) all);
    }
    free(all); free(keep);
    return out;
}
