/*
 * apss_jni.c -- JNI shim between the reference's JVM (Scala 2.10 / Akka 2.3.4) and include/apss.h.
 *
 * UNVERIFIED SOURCE: this image has no JDK (no jni.h, no javac/scalac), so this file has never been
 * compiled.  It is what a maintainer adds next to integration/scala/GpuIndexingWorkerActor.scala;
 * build with   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *                  apss_jni.c -L../../all-pairs-similarity_b200 -lapss_b200 -o libapss_jni.so
 *
 * Java side (cpslab.gpu.ApssNative, see the Scala file):
 *   static native long   create(int dim, double simThr, double idxThr, int device, int semantics);
 *   static native void   destroy(long h);
 *   static native long[] insertBatch(long h, long[] indptr, int[] indices, double[] values,
 *                                    long[] extKeys, int[] firstDim, int flags);   // -> apss_batch_result as longs
 *   static native int    fetchPairs(long h, int[] q, int[] c, double[] sim);
 *   static native void   fetchStatus(long h, byte[] status);
 *   static native void   freeze(long h);
 * Errors become java.lang.RuntimeException, which the actor catches and logs exactly like
 * IndexingWorkerActor.scala:135-137 (batch dropped, index unchanged).
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>
#include "apss.h"

static void throw_rt(JNIEnv *env, apss_handle *h, int32_t rc) {
  char msg[600];
  snprintf(msg, sizeof msg, "apss error %d: %s", rc, h ? apss_last_error(h) : "no handle");
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
}

JNIEXPORT jlong JNICALL Java_cpslab_gpu_ApssNative_create(JNIEnv *env, jclass cls, jint dim, jdouble sim_thr,
                                                          jdouble idx_thr, jint device, jint semantics, jint pruning) {
  apss_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (int32_t)sizeof cfg;
  cfg.dim = dim; cfg.similarity_threshold = sim_thr; cfg.index_threshold = idx_thr;
  cfg.device = device; cfg.semantics = semantics;
  cfg.pruning = pruning;   /* 0 parity counters, 2 exact index reduction (same pairs, far less work); see include/apss.h */
  apss_handle *h = NULL;
  int32_t rc = apss_create(&cfg, &h);
  if (rc != APSS_OK) { throw_rt(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_cpslab_gpu_ApssNative_destroy(JNIEnv *env, jclass cls, jlong h) {
  apss_destroy((apss_handle *)(intptr_t)h);
}

JNIEXPORT jlongArray JNICALL Java_cpslab_gpu_ApssNative_insertBatch(JNIEnv *env, jclass cls, jlong hh, jlongArray indptr,
                                                                    jintArray indices, jdoubleArray values, jlongArray ext_keys,
                                                                    jintArray first_dim, jint flags) {
  apss_handle *h = (apss_handle *)(intptr_t)hh;
  const jsize n = (*env)->GetArrayLength(env, indptr) - 1;
  /* borrowed for the duration of the call; the library copies to the device (SURVEY 8(b) "ownership") */
  jlong *ip = (*env)->GetPrimitiveArrayCritical(env, indptr, NULL);
  jint *ix = (*env)->GetPrimitiveArrayCritical(env, indices, NULL);
  jdouble *vv = (*env)->GetPrimitiveArrayCritical(env, values, NULL);
  jlong *ek = ext_keys ? (*env)->GetPrimitiveArrayCritical(env, ext_keys, NULL) : NULL;
  jint *fd = first_dim ? (*env)->GetPrimitiveArrayCritical(env, first_dim, NULL) : NULL;
  apss_batch_result res;
  int32_t rc = apss_insert_batch(h, (int32_t)n, (const int64_t *)ip, (const int32_t *)ix, (const double *)vv,
                                 (const int64_t *)ek, (const int32_t *)fd, (uint32_t)flags, &res);
  if (fd) (*env)->ReleasePrimitiveArrayCritical(env, first_dim, fd, JNI_ABORT);
  if (ek) (*env)->ReleasePrimitiveArrayCritical(env, ext_keys, ek, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, values, vv, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, indices, ix, JNI_ABORT);
  (*env)->ReleasePrimitiveArrayCritical(env, indptr, ip, JNI_ABORT);
  if (rc != APSS_OK) { throw_rt(env, h, rc); return NULL; }
  jlong out[8] = {res.id_base, res.n_pairs, res.n_rejected, res.n_empty, res.n_active,
                  res.postings_visited, res.candidates_unique, res.n_prefilter};
  jlongArray arr = (*env)->NewLongArray(env, 8);
  (*env)->SetLongArrayRegion(env, arr, 0, 8, out);
  return arr;
}

JNIEXPORT jint JNICALL Java_cpslab_gpu_ApssNative_fetchPairs(JNIEnv *env, jclass cls, jlong hh, jintArray q, jintArray c,
                                                             jdoubleArray sim) {
  apss_handle *h = (apss_handle *)(intptr_t)hh;
  const jsize cap = (*env)->GetArrayLength(env, q);
  jint *qq = (*env)->GetPrimitiveArrayCritical(env, q, NULL);
  jint *cc = (*env)->GetPrimitiveArrayCritical(env, c, NULL);
  jdouble *ss = (*env)->GetPrimitiveArrayCritical(env, sim, NULL);
  int64_t n = 0;
  int32_t rc = apss_fetch_pairs(h, (int32_t *)qq, (int32_t *)cc, (double *)ss, cap, &n);
  (*env)->ReleasePrimitiveArrayCritical(env, sim, ss, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, c, cc, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, q, qq, 0);
  if (rc != APSS_OK) { throw_rt(env, h, rc); return 0; }
  return (jint)n;
}

JNIEXPORT void JNICALL Java_cpslab_gpu_ApssNative_fetchStatus(JNIEnv *env, jclass cls, jlong hh, jbyteArray status) {
  apss_handle *h = (apss_handle *)(intptr_t)hh;
  const jsize cap = (*env)->GetArrayLength(env, status);
  jbyte *st = (*env)->GetPrimitiveArrayCritical(env, status, NULL);
  int32_t rc = apss_fetch_status(h, (uint8_t *)st, cap);
  (*env)->ReleasePrimitiveArrayCritical(env, status, st, 0);
  if (rc != APSS_OK) throw_rt(env, h, rc);
}

JNIEXPORT void JNICALL Java_cpslab_gpu_ApssNative_freeze(JNIEnv *env, jclass cls, jlong hh) {
  apss_freeze((apss_handle *)(intptr_t)hh);
}
