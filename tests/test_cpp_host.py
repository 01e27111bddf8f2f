"""The C++ host side above the C ABI (all-pairs-similarity_b200/host/apss_actor.hpp): GpuIndexingWorkerActor,
RegionRouter and ClientConnection driven like the reference's actors.  tests/cpp/actor_scenario.cpp is built with
g++ and run: on CPU against the oracle-backed test double (tests/cpp/oracle_engine.hpp), on the GPU against
libapss_b200.so through include/apss.h."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "actor_scenario.cpp")
INC = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "all-pairs-similarity_b200", "host"),
       "-I" + os.path.join(ROOT, "tests", "cpp")]
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _cuda_lib_dirs():
    """where libcudart.so.12 (a dependency of libapss_b200.so) lives: the directories of the CUDA runtime this
    process has loaded through torch, plus the toolkit's"""
    dirs = []
    try:
        import torch  # noqa: F401  (loads libcudart)
        with open("/proc/self/maps") as f:
            for ln in f:
                if "libcudart" in ln:
                    d = os.path.dirname(ln.split()[-1])
                    if d not in dirs:
                        dirs.append(d)
    except Exception:
        pass
    for d in ("/usr/local/cuda/lib64",):
        if os.path.isdir(d) and d not in dirs:
            dirs.append(d)
    return dirs


def _env(extra_dirs):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(list(extra_dirs) + [p for p in env.get("LD_LIBRARY_PATH", "").split(":") if p])
    return env


def _build(tmp_path, name, lib, defines=(), lib_dirs=()):
    exe = str(tmp_path / name)
    cmd = [GXX, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", *defines, *INC, SRC, "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    for d in lib_dirs:
        cmd += ["-Wl,-rpath-link," + d, "-Wl,-rpath," + d]
    r = subprocess.run(cmd, capture_output=True, text=True, env=_env(lib_dirs))
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


def test_cpp_actor_against_the_oracle_double(tmp_path):
    from oracle import oracle as orc
    exe = _build(tmp_path, "actor_cpu", orc.build(), ["-DUSE_ORACLE_ENGINE"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "actor scenario ok" in r.stdout
    assert "vector1 size: 32, vector2 size: 64" in r.stderr          # the swallowed exception is logged (IWA:135-137)


def _device_lists():
    import torch
    out = [[], [0, 0]]                       # one GPU; two shards on GPU 0 (test hook: the dispatch logic on a one-GPU box)
    if torch.cuda.is_available() and torch.cuda.device_count() >= 2:
        out.append(list(range(torch.cuda.device_count())))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("pruning", [0, 2, 3])
def test_cpp_actor_through_the_c_abi(tmp_path, pruning):
    """the C++ actor holds ONE engine; with device_ids = {0..N-1} the shard dispatch happens below the C ABI"""
    import apss_b200
    lib = apss_b200.native.LIB_PATH
    assert os.path.exists(lib), "libapss_b200.so has not been built"
    dirs = _cuda_lib_dirs()
    exe = _build(tmp_path, "actor_gpu", lib, lib_dirs=dirs)
    for devs in _device_lists():
        env = _env(dirs)
        env["APSS_TEST_ALLOW_DUP_DEVICES"] = "1"
        r = subprocess.run([exe, str(pruning)] + [str(d) for d in devs], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "actor scenario ok (pruning=%d, devices=%d)" % (pruning, max(1, len(devs))) in r.stdout
