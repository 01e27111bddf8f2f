/*
 * apss_jni.c -- JNI shim between the reference's JVM (Scala 2.10 / Akka 2.3.4) and include/apss.h.
 *
 * Java side: integration/java/cpslab/gpu/ApssNative.java (a Java class with static natives in package
 * cpslab.gpu, hence the symbol names Java_cpslab_gpu_ApssNative_* with a jclass receiver below).
 *
 * UNVERIFIED ON A JVM: this image has no JDK.  The file is compiled against a minimal stand-in jni.h
 * (tests/stubs/jni.h) by tests/test_integration_sources.py, which also checks that every native method
 * the Java class declares has its symbol here.  Real build:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       apss_jni.c -L../../all-pairs-similarity_b200 -lapss_b200 -o libapss_jni.so
 *
 * Arrays are COPIED out of the Java heap (Get<T>ArrayRegion) before the library is called:
 * apss_insert_batch blocks on CUDA stream synchronisation, and a JNI critical region must not be held
 * across a blocking call (it stalls the collector).  Errors become java.lang.RuntimeException, which the
 * actor catches and logs exactly like IndexingWorkerActor.scala:135-137 (batch dropped, index unchanged).
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "apss.h"

static void throw_rt(JNIEnv *env, apss_handle *h, int32_t rc, const char *what) {
  char msg[700];
  snprintf(msg, sizeof msg, "apss error %d: %s", (int)rc, what ? what : (h ? apss_last_error(h) : "no handle"));
  jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
  if (cls) (*env)->ThrowNew(env, cls, msg);
}

JNIEXPORT jlong JNICALL Java_cpslab_gpu_ApssNative_create(JNIEnv *env, jclass cls, jint dim, jdouble sim_thr, jdouble idx_thr,
                                                          jintArray device_ids, jint semantics, jint pruning) {
  apss_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_size = (int32_t)sizeof cfg;
  cfg.dim = dim; cfg.similarity_threshold = sim_thr; cfg.index_threshold = idx_thr;
  cfg.semantics = semantics;   /* APSS_SEM_R1 / APSS_SEM_R0 (as built: needs firstDim[] per batch) */
  cfg.pruning = pruning;       /* 0 parity counters; 3 exact index reduction, same pairs (include/apss.h) */
  jsize nd = device_ids ? (*env)->GetArrayLength(env, device_ids) : 0;
  if (nd > APSS_MAX_DEVICES) { throw_rt(env, NULL, APSS_E_INVALID, "too many devices"); return 0; }
  if (nd > 0) {
    jint ids[APSS_MAX_DEVICES];
    (*env)->GetIntArrayRegion(env, device_ids, 0, nd, ids);
    cfg.device = ids[0]; cfg.n_devices = nd;
    for (jsize i = 0; i < nd; i++) cfg.device_ids[i] = ids[i];
  }
  apss_handle *h = NULL;
  int32_t rc = apss_create(&cfg, &h);
  if (rc != APSS_OK) { throw_rt(env, NULL, rc, "apss_create failed"); return 0; }
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
  const jsize nnz = (*env)->GetArrayLength(env, indices);
  if (n < 0 || (*env)->GetArrayLength(env, values) != nnz) { throw_rt(env, h, APSS_E_INVALID, "ragged batch arrays"); return NULL; }
  /* one native block: indptr | values | ext_keys (8-byte parts first), then indices | first_dim */
  const size_t b_ptr = sizeof(jlong) * (size_t)(n + 1), b_val = sizeof(jdouble) * (size_t)nnz;
  const size_t b_key = ext_keys ? sizeof(jlong) * (size_t)n : 0, b_idx = sizeof(jint) * (size_t)nnz;
  const size_t b_fd = first_dim ? sizeof(jint) * (size_t)n : 0;
  char *blk = (char *)malloc(b_ptr + b_val + b_key + b_idx + b_fd + 8);
  if (!blk) { throw_rt(env, h, APSS_E_NOMEM, "out of native memory"); return NULL; }
  jlong *ip = (jlong *)blk; jdouble *vv = (jdouble *)(blk + b_ptr); jlong *ek = ext_keys ? (jlong *)(blk + b_ptr + b_val) : NULL;
  jint *ix = (jint *)(blk + b_ptr + b_val + b_key); jint *fd = first_dim ? (jint *)(blk + b_ptr + b_val + b_key + b_idx) : NULL;
  (*env)->GetLongArrayRegion(env, indptr, 0, n + 1, ip);
  (*env)->GetIntArrayRegion(env, indices, 0, nnz, ix);
  (*env)->GetDoubleArrayRegion(env, values, 0, nnz, vv);
  if (ek) (*env)->GetLongArrayRegion(env, ext_keys, 0, n, ek);
  if (fd) (*env)->GetIntArrayRegion(env, first_dim, 0, n, fd);
  apss_batch_result res;
  int32_t rc = apss_insert_batch(h, (int32_t)n, (const int64_t *)ip, (const int32_t *)ix, (const double *)vv,
                                 (const int64_t *)ek, (const int32_t *)fd, (uint32_t)flags, &res);
  free(blk);
  if (rc != APSS_OK) { throw_rt(env, h, rc, NULL); return NULL; }
  jlong out[8] = {res.id_base, res.n_pairs, res.n_rejected, res.n_empty, res.n_active,
                  res.postings_visited, res.candidates_unique, res.n_prefilter};
  jlongArray arr = (*env)->NewLongArray(env, 8);
  if (arr) (*env)->SetLongArrayRegion(env, arr, 0, 8, out);
  return arr;
}

JNIEXPORT jint JNICALL Java_cpslab_gpu_ApssNative_fetchPairs(JNIEnv *env, jclass cls, jlong hh, jintArray q, jintArray c,
                                                             jdoubleArray sim) {
  apss_handle *h = (apss_handle *)(intptr_t)hh;
  const jsize cap = (*env)->GetArrayLength(env, q);
  if ((*env)->GetArrayLength(env, c) < cap || (*env)->GetArrayLength(env, sim) < cap) { throw_rt(env, h, APSS_E_INVALID, "output arrays differ in length"); return 0; }
  char *blk = (char *)malloc((sizeof(jdouble) + 2 * sizeof(jint)) * (size_t)(cap > 0 ? cap : 1));
  if (!blk) { throw_rt(env, h, APSS_E_NOMEM, "out of native memory"); return 0; }
  jdouble *ss = (jdouble *)blk; jint *qq = (jint *)(blk + sizeof(jdouble) * (size_t)cap); jint *cc = qq + cap;
  int64_t n = 0;
  int32_t rc = apss_fetch_pairs(h, (int32_t *)qq, (int32_t *)cc, (double *)ss, cap, &n);
  if (rc == APSS_OK) {
    const jsize m = (jsize)(n < cap ? n : cap);
    (*env)->SetIntArrayRegion(env, q, 0, m, qq);
    (*env)->SetIntArrayRegion(env, c, 0, m, cc);
    (*env)->SetDoubleArrayRegion(env, sim, 0, m, ss);
  }
  free(blk);
  if (rc != APSS_OK) { throw_rt(env, h, rc, NULL); return 0; }
  return (jint)n;
}

JNIEXPORT void JNICALL Java_cpslab_gpu_ApssNative_fetchStatus(JNIEnv *env, jclass cls, jlong hh, jbyteArray status) {
  apss_handle *h = (apss_handle *)(intptr_t)hh;
  const jsize cap = (*env)->GetArrayLength(env, status);
  jbyte *st = (jbyte *)malloc((size_t)(cap > 0 ? cap : 1));
  if (!st) { throw_rt(env, h, APSS_E_NOMEM, "out of native memory"); return; }
  int32_t rc = apss_fetch_status(h, (uint8_t *)st, cap);
  if (rc == APSS_OK) (*env)->SetByteArrayRegion(env, status, 0, cap, st);
  free(st);
  if (rc != APSS_OK) throw_rt(env, h, rc, NULL);
}

JNIEXPORT void JNICALL Java_cpslab_gpu_ApssNative_freeze(JNIEnv *env, jclass cls, jlong hh) {
  apss_freeze((apss_handle *)(intptr_t)hh);
}
