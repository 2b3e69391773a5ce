"""AddressSanitizer + UBSan runs of the two pieces of host C/C++ that tests and bench legs rely on: the C
oracle (its rehash bug of round 1 was a heap-use-after-free that ASan reports on the first growth) and the
native matcher.  Built and run as stand-alone programs (tests/san/); skipped if the compiler has no sanitizer
runtime."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAN = ["-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"]


def build_and_run(tmp_path, compiler, sources, includes, extra=()):
    cc = shutil.which(compiler)
    if cc is None:
        pytest.skip(f"{compiler} not found")
    exe = str(tmp_path / "san_prog")
    cmd = [cc, *SAN, *[f"-I{i}" for i in includes], *sources, "-lpthread", *extra, "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and ("sanitize" in b.stderr or "asan" in b.stderr.lower()):
        pytest.skip("no sanitizer runtime for this compiler")
    assert b.returncode == 0, b.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, (r.stdout + r.stderr)[-3000:]
    return r.stdout


def test_oracle_is_clean_under_asan_and_ubsan(tmp_path):
    out = build_and_run(tmp_path, "gcc", [os.path.join(ROOT, "tests", "san", "san_oracle.c"),
                                          os.path.join(ROOT, "oracle", "limitador_oracle.c")],
                        [os.path.join(ROOT, "oracle")])
    assert out.startswith("ok limited=")


def test_matcher_is_clean_under_asan_and_ubsan(tmp_path):
    out = build_and_run(tmp_path, "g++", [os.path.join(ROOT, "tests", "san", "san_matcher.cpp"),
                                          os.path.join(ROOT, "limitador_b200", "csrc", "rl_match.cpp")],
                        [os.path.join(ROOT, "include")], extra=("-std=c++17",))
    assert out.startswith("ok added=")


def test_rls_wire_surface_is_clean_under_asan_and_ubsan(tmp_path):
    """The decoder reads bytes from the network: 200 000 mutated / truncated / random messages, then the plan and finish
    stages over batches mixing good, malformed and unsupported requests, 1 and 3 workers."""
    csrc = os.path.join(ROOT, "limitador_b200", "csrc")
    out = build_and_run(tmp_path, "g++", [os.path.join(ROOT, "tests", "san", "san_rls.cpp"), os.path.join(csrc, "rl_rls.cpp"),
                                          os.path.join(csrc, "rl_match.cpp")],
                        [os.path.join(ROOT, "include")], extra=("-std=c++17",))
    assert out.startswith("ok decoded=")


def test_crdt_oracle_and_shim_kernels_are_clean_under_asan_and_ubsan(tmp_path):
    """The maintenance / CRDT kernels under the host shim with ASan + UBSan: out-of-bounds row arithmetic in a kernel shows
    up here, in the container without a GPU (tests/san/san_kernels.cpp)."""
    out = build_and_run(tmp_path, "g++", [os.path.join(ROOT, "tests", "san", "san_kernels.cpp")],
                        [os.path.join(ROOT, "include")], extra=("-std=c++17",))
    assert out.startswith("ok kernels")


@pytest.mark.parametrize("flags", [SAN, ["-g", "-O1", "-fsanitize=thread"]], ids=["asan+ubsan", "tsan"])
def test_front_with_the_matcher_inside_is_clean_under_the_sanitizers(tmp_path, flags):
    """rl_front_check_and_update_bindings from 8 threads while limits are added and deleted: address / UB and data races
    (the store call is a stub that answers from the CSR the dispatcher built)."""
    import shutil as _sh
    cc = _sh.which("g++")
    if cc is None:
        pytest.skip("g++ not found")
    csrc = os.path.join(ROOT, "limitador_b200", "csrc")
    exe = str(tmp_path / "san_front")
    cmd = [cc, *flags, "-std=c++17", f"-I{os.path.join(ROOT, 'include')}", os.path.join(ROOT, "tests", "san", "san_front.cpp"),
           "-x", "c++", os.path.join(csrc, "rl_front.cu"), "-x", "none", os.path.join(csrc, "rl_match.cpp"), "-lpthread", "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and ("sanitize" in b.stderr or "tsan" in b.stderr.lower() or "asan" in b.stderr.lower()):
        pytest.skip("no sanitizer runtime for this compiler")
    assert b.returncode == 0, b.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1", TSAN_OPTIONS="halt_on_error=1")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert r.returncode == 0 and "ERROR" not in r.stderr and "WARNING: ThreadSanitizer" not in r.stderr and "runtime error" not in r.stderr, \
        (r.stdout + r.stderr)[-3000:]
    assert r.stdout.startswith("ok front")
