"""Bounds check of the kernel sources on the CPU: a subset of the emulator tests under the AddressSanitizer build of
tests/emu (every emulated global / shared-memory access is checked).  The whole emulator suite runs clean under it too
(4 min; command in tests/emu/Makefile); this keeps a fast slice in the default CPU run.  compute-sanitizer is closed on the
GPU pool, so this is the memory-safety evidence for the index arithmetic."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUBSET = ("test_split_plans or test_bluestein_padded or test_mixed_radix_many or test_device_resident or test_strided_plan "
          "or test_batch_not_multiple or test_clenshaw_tile or test_sample2d_fused or test_empty_and_vector")


def test_emulated_kernels_under_address_sanitizer():
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not found")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "libafb_emu_asan.so"], stdout=subprocess.DEVNULL)
    env = dict(os.environ, AFB_EMU_SO=os.path.join(ROOT, "tests", "emu", "libafb_emu_asan.so"), LD_PRELOAD=asan,
               ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_emulated_kernels.py"), "-q", "-x",
                        "-p", "no:cacheprovider", "-k", SUBSET], env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "AddressSanitizer" not in tail, tail
