// ApssNative.java -- JVM side of integration/jni/apss_jni.c (include/apss.h).
//
// A plain Java class with STATIC natives in package cpslab.gpu, so that the JNI symbol names are
// Java_cpslab_gpu_ApssNative_<method>(JNIEnv*, jclass, ...) -- exactly what apss_jni.c exports
// (tests/test_integration_sources.py checks the two lists against each other).  sbt compiles
// src/main/java next to src/main/scala; the Scala actor calls cpslab.gpu.ApssNative.create(...).
//
// UNVERIFIED ON A JVM: this image has no JDK.
package cpslab.gpu;

public final class ApssNative {
  static { System.loadLibrary("apss_jni"); }
  private ApssNative() {}

  public static final int SEM_R1 = 0, SEM_R0 = 1;                 // include/apss.h APSS_SEM_*
  public static final int QUERY_ONLY = 1, SKIP_ADMIT = 4;         // APSS_BATCH_*

  /** apss_create.  deviceIds: the GPUs that share the index (id-range shards below the C ABI); one entry = one GPU. */
  public static native long create(int dim, double simThr, double idxThr, int[] deviceIds, int semantics, int pruning);
  public static native void destroy(long h);
  /** apss_insert_batch; returns {id_base, n_pairs, n_rejected, n_empty, n_active, postings_visited, candidates_unique, n_prefilter}. */
  public static native long[] insertBatch(long h, long[] indptr, int[] indices, double[] values,
                                          long[] extKeys, int[] firstDim, int flags);
  public static native int fetchPairs(long h, int[] q, int[] c, double[] sim);
  public static native void fetchStatus(long h, byte[] status);
  public static native void freeze(long h);
}
