/* jni_stub.h — COMPILE-CHECK STAND-IN for the JDK's <jni.h>.
 *
 * This image has no JDK, so bfq_gpumatch_jni.c cannot be built against the real header here. This file declares just the
 * JNI types and JNIEnv entry points the shim uses, with the signatures the JNI specification gives them, so that
 * `gcc -fsyntax-only -DBFQ_JNI_STUB` (run by __graft_entry__.build()) type-checks every call the shim makes into
 * include/bfq_gpumatch.h. It is NOT ABI-compatible with a JVM (the function table is not in JVM order): a maintainer builds the
 * shim WITHOUT -DBFQ_JNI_STUB against $JAVA_HOME/include/jni.h.
 */
#ifndef BFQ_JNI_STUB_H
#define BFQ_JNI_STUB_H
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jbyteArray;
typedef jarray jobjectArray;
typedef jobject jthrowable;

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
    void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
    jlong (*GetDirectBufferCapacity)(JNIEnv* env, jobject buf);
    jobject (*NewDirectByteBuffer)(JNIEnv* env, void* address, jlong capacity);
    jsize (*GetArrayLength)(JNIEnv* env, jarray array);
    jint* (*GetIntArrayElements)(JNIEnv* env, jintArray array, jboolean* isCopy);
    void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray array, jint* elems, jint mode);
    jlong* (*GetLongArrayElements)(JNIEnv* env, jlongArray array, jboolean* isCopy);
    void (*ReleaseLongArrayElements)(JNIEnv* env, jlongArray array, jlong* elems, jint mode);
    jbyte* (*GetByteArrayElements)(JNIEnv* env, jbyteArray array, jboolean* isCopy);
    void (*ReleaseByteArrayElements)(JNIEnv* env, jbyteArray array, jbyte* elems, jint mode);
    jlongArray (*NewLongArray)(JNIEnv* env, jsize len);
    void (*SetLongArrayRegion)(JNIEnv* env, jlongArray array, jsize start, jsize len, const jlong* buf);
    jbyteArray (*NewByteArray)(JNIEnv* env, jsize len);
    void (*SetByteArrayRegion)(JNIEnv* env, jbyteArray array, jsize start, jsize len, const jbyte* buf);
    jobjectArray (*NewObjectArray)(JNIEnv* env, jsize len, jclass clazz, jobject init);
    void (*SetObjectArrayElement)(JNIEnv* env, jobjectArray array, jsize index, jobject val);
};
#endif
