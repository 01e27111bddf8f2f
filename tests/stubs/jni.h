/* Minimal stand-in for the JDK's jni.h: just enough of the JNI 1.6 surface for integration/jni/apss_jni.c to be
 * parsed and compiled in an image without a JDK (tests/test_integration_sources.py).  Types and the member
 * names / signatures of JNINativeInterface_ follow the JNI specification; the table is NOT layout-compatible with a
 * real JVM and nothing built against this file may be loaded into one. */
#ifndef APSS_STUB_JNI_H_
#define APSS_STUB_JNI_H_
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef double jdouble; typedef jint jsize;
struct _jobject; typedef struct _jobject *jobject;
typedef jobject jclass; typedef jobject jthrowable; typedef jobject jarray;
typedef jarray jintArray; typedef jarray jlongArray; typedef jarray jdoubleArray; typedef jarray jbyteArray;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv *, const char *);
  jint (*ThrowNew)(JNIEnv *, jclass, const char *);
  jsize (*GetArrayLength)(JNIEnv *, jarray);
  jlongArray (*NewLongArray)(JNIEnv *, jsize);
  void (*GetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, jint *);
  void (*GetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, jlong *);
  void (*GetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, jdouble *);
  void (*SetIntArrayRegion)(JNIEnv *, jintArray, jsize, jsize, const jint *);
  void (*SetLongArrayRegion)(JNIEnv *, jlongArray, jsize, jsize, const jlong *);
  void (*SetDoubleArrayRegion)(JNIEnv *, jdoubleArray, jsize, jsize, const jdouble *);
  void (*SetByteArrayRegion)(JNIEnv *, jbyteArray, jsize, jsize, const jbyte *);
};
#endif
