"""The oracle is the checker of every parity claim: its own tests (known answers, properties, golden vectors) are run once more
against a build of the same C sources with AddressSanitizer + UBSan, in a child process with the sanitizer runtimes preloaded."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_tests_pass_under_address_and_ub_sanitizers():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "asan"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    libs = [subprocess.run([cc, f"-print-file-name={n}"], capture_output=True, text=True).stdout.strip() for n in ("libasan.so", "libubsan.so")]
    assert all(os.path.isabs(p) for p in libs), libs
    env = dict(os.environ, DMK_ORACLE_LIBRARY=os.path.join(ROOT, "oracle", "liboracle_asan.so"), LD_PRELOAD=":".join(libs),
               ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_oracle_cull.py"), os.path.join(ROOT, "tests", "test_oracle_pose.py"),
                        os.path.join(ROOT, "tests", "test_oracle_sampling.py"), os.path.join(ROOT, "tests", "test_golden.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
