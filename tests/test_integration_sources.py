"""The reference-side binding cannot be run here -- no JVM -- but it can be kept honest: the JNI shim
must compile (against a stand-in jni.h that follows the JNI specification's signatures), export exactly one symbol per
native method of the Java class the Scala actor calls, with the mangled names a JVM will look for, and never hold a
JNI critical region across the blocking library call."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI_C = os.path.join(ROOT, "integration", "jni", "apss_jni.c")
JAVA = os.path.join(ROOT, "integration", "java", "cpslab", "gpu", "ApssNative.java")
SCALA = os.path.join(ROOT, "integration", "scala", "GpuIndexingWorkerActor.scala")


def _java_natives():
    src = open(JAVA).read()
    pkg = re.search(r"^package\s+([\w.]+);", src, re.M).group(1)
    cls = re.search(r"public final class (\w+)", src).group(1)
    flat = " ".join(src.split())
    natives = re.findall(r"public static native [\w\[\]]+ (\w+)\(([^)]*)\)", flat)
    return pkg, cls, {name: [a for a in args.split(",") if a.strip()] for name, args in natives}


def test_jni_shim_compiles_and_exports_the_symbols_the_jvm_will_look_for(tmp_path):
    obj = tmp_path / "apss_jni.o"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-fPIC", "-c", JNI_C,
                           "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), "-o", str(obj)])
    syms = {l.split()[-1] for l in subprocess.check_output(["nm", "-g", "--defined-only", str(obj)], text=True).splitlines() if " T " in l}
    pkg, cls, natives = _java_natives()
    assert pkg == "cpslab.gpu" and cls == "ApssNative" and len(natives) == 6
    want = {"Java_%s_%s_%s" % (pkg.replace(".", "_"), cls, m) for m in natives}
    assert syms == want, (sorted(syms), sorted(want))
    # static natives take (JNIEnv*, jclass, args...): the C parameter count must be the Java one + 2
    csrc = " ".join(open(JNI_C).read().split())
    for m, jargs in natives.items():
        cargs = re.search(r"Java_cpslab_gpu_ApssNative_%s\(([^)]*)\)" % m, csrc).group(1).split(",")
        assert len(cargs) == len(jargs) + 2 and "jclass" in cargs[1], (m, cargs, jargs)


def test_jni_shim_holds_no_critical_region_and_scala_side_matches():
    c = open(JNI_C).read()
    body = c[c.index("#include <jni.h>"):]            # the header comment may name the call it avoids
    assert "GetPrimitiveArrayCritical" not in body and "#include <stdio.h>" in c
    sc = open(SCALA).read()
    assert "import cpslab.gpu.ApssNative" in sc and "object ApssNative" not in sc
    _, _, natives = _java_natives()
    for m in re.findall(r"ApssNative\.(\w+)\(", sc):
        assert m in natives, m
    # the as-built semantics are reachable from the actor: config key, firstDim from the wrapper's own Set
    assert "cpslab.allpair.gpu.semantics" in sc and "w.indices.head" in sc and "ApssNative.SEM_R0" in sc
    # ids are recorded only after the native call returned (a refused batch leaves no trace)
    assert sc.index("ApssNative.insertBatch(") < sc.index("firstOf ++= fresh")
