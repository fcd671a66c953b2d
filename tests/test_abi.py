"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and fails loudly (AFB_NO_DEVICE) instead of falling back when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    import approxfun_b200 as af

    return af


def test_header_and_library_symbols_agree(built):
    hdr = open(os.path.join(ROOT, "include", "approxfun_b200.h")).read()
    declared = sorted(set(re.findall(r"^int32_t\s+(afb_\w+)\s*\(", hdr, flags=re.M)))
    assert declared == sorted(built._lib.SYMBOLS)
    lib = ctypes.CDLL(built.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.afb_version() == 100


def test_every_entry_point_cites_the_reference():
    hdr = open(os.path.join(ROOT, "include", "approxfun_b200.h")).read()
    for needle in ("src/Extras/sample.jl:55-75", "src/Extras/sample.jl:170-175", "src/Extras/fftGeneric.jl:47-54",
                   "ext/ApproxFunDualNumbersExt.jl:10-12", "NEWS.md:92", "docs/src/usage/spaces.md:113-125"):
        assert needle in hdr


def test_no_cpu_fallback_without_device(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = built._lib.load()
    h = ctypes.c_void_p()
    st = lib.afb_plan_create(ctypes.byref(h), built._lib.CHEB1_FWD, 64, 0, 1, 64, 0)
    assert st == built._lib.AFB_NO_DEVICE
    buf = ctypes.create_string_buffer(256)
    assert lib.afb_last_error(buf, 256) == built._lib.AFB_NO_DEVICE
    assert b"no CPU fallback" in buf.value
    with pytest.raises(built.AfbError):
        built.transform(built.Chebyshev(), np.zeros(64))
    with pytest.raises(built.AfbError):
        built.clenshaw(built.Chebyshev(), np.ones(3), 0.5)


def test_round2_entry_points_fail_loudly_without_device(built):
    # the multi-device init, device buffers, piecewise bisection, mixed tensor plans and the DMMA variant: no GPU => AFB_NO_DEVICE
    # (or an argument error that is raised before the device is touched), never a CPU result
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = built._lib
    lib = L.load()
    ND = L.AFB_NO_DEVICE
    devs = (ctypes.c_int32 * 2)(0, 1)
    assert lib.afb_init(devs, 2) == ND and lib.afb_init(None, 0) == ND
    h = ctypes.c_void_p()
    assert lib.afb_buf_create(ctypes.byref(h), 16) == ND and not h.value
    for kind in (L.FOURIER_CHEB1_2D_FWD, L.LAURENT_CHEB1_2D_INV):
        assert lib.afb_plan_create(ctypes.byref(h), kind, 8, 8, 1, 8, 0) == ND
    assert lib.afb_plan_create(ctypes.byref(h), L.CHEB1_FWD, 1000, 0, 4, 1000, 0) == ND          # mixed-radix length
    c = np.ones(4); offs = np.array([0, 4], dtype=np.int64); brk = np.array([-1.0, 1.0]); u = np.full(3, 0.5); out = np.empty(3)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert lib.afb_piecewise_bisect(p(c), p(offs), p(brk), 1, p(u), p(out), 3, 47) == ND
    assert lib.afb_piecewise_clenshaw(p(c), p(offs), p(brk), 1, p(u), p(out), 3) == ND
    assert lib.afb_clenshaw_fourier_cheb2d(p(c), 2, 2, p(u), p(u), p(out), 3) == ND
    assert lib.afb_dgemm_nn_dmma_device(p(c), p(c), p(out), 1, 1, 1, None) == ND
    assert lib.afb_host_register(p(c), c.nbytes) == ND and lib.afb_host_unregister(p(c)) == ND
    # argument errors come first
    assert lib.afb_host_register(None, 8) == L.AFB_BAD_ARG and lib.afb_host_register(p(c), 0) == L.AFB_BAD_ARG
    assert lib.afb_init(None, 2) == L.AFB_BAD_ARG
    assert lib.afb_piecewise_bisect(p(c), p(offs), p(np.array([1.0, -1.0])), 1, p(u), p(out), 3, 47) == L.AFB_BAD_SIZE
    assert lib.afb_buf_create(ctypes.byref(h), -1) == L.AFB_BAD_SIZE
    assert lib.afb_sharded2d_set_chunks(None, 2) == L.AFB_BAD_ARG and lib.afb_sharded2d_set_tuning(None, 0, 8, 0) == L.AFB_BAD_ARG
    with pytest.raises(built.AfbError):
        built.device.DeviceArray(8)
    with pytest.raises(built.AfbError):
        built.api.use_devices([0])


def test_argument_validation_needs_no_device(built):
    lib = built._lib.load()
    h = ctypes.c_void_p()
    assert lib.afb_plan_create(None, 0, 8, 0, 1, 8, 0) == built._lib.AFB_BAD_ARG
    assert lib.afb_plan_create(ctypes.byref(h), 99, 8, 0, 1, 8, 0) == built._lib.AFB_BAD_ARG
    assert lib.afb_plan_create(ctypes.byref(h), 0, -1, 0, 1, 8, 0) == built._lib.AFB_BAD_SIZE
    assert lib.afb_clenshaw_cheb(None, -1, None, None, 0) == built._lib.AFB_BAD_SIZE
    assert lib.afb_cheb_bisect_cols(None, 4, 2, None, None, 3, 47) == built._lib.AFB_BAD_SIZE  # ld < N
    out = np.empty(4)
    assert lib.afb_fourierpoints(out.ctypes.data, 4, 0.0, 1.0) == 0
    assert np.array_equal(out, [0, 0.25, 0.5, 0.75])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "approxfun.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt and "libafb_emu" not in txt, f


def test_golden_fixture_matches_oracle():
    from oracle import approxfun_oracle as orc
    from oracle import oracle_c as orcc

    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    for n in (8, 64, 100, 4096):
        assert np.max(np.abs(orc.cheb_transform(g[f"v_{n}"]) - g[f"cheb_fwd_{n}"])) < 1e-15
        assert np.max(np.abs(orc.fourier_itransform(g[f"v_{n}"]) - g[f"fourier_inv_{n}"])) < 1e-13
    assert np.array_equal(orcc.clenshaw(g["cl_c"], g["cl_x"]), g["cl_y"])
    assert np.array_equal(orcc.cheb_bisect(g["bis_c"], g["bis_u"]), g["bis_x"])
    assert np.array_equal(orcc.cheb_normcumsum(g["ncs_in"]), g["ncs_out"])


def _build_c_smoke(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(root, "approxfun.jl_b200")
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(root, "tests", "c", "abi_smoke.c"),
                           "-L" + libdir, "-l:libapproxfun_b200.so", "-lm", "-Wl,-rpath," + libdir])
    return exe


def test_plain_c_caller_links_and_fails_loudly_without_a_gpu(tmp_path):
    # a C host (or Julia's ccall) against include/approxfun_b200.h: compiles as C, links, and without a device gets
    # AFB_NO_DEVICE -- never a CPU result
    import subprocess
    import torch
    exe = _build_c_smoke(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "abi smoke ok" in out.stdout, out.stderr
    else:
        assert out.returncode == 3 and "no CUDA device" in out.stderr, (out.returncode, out.stderr)


def test_host_side_index_helpers():
    # pure host logic of the Python mirror (no device): total-degree interlace <-> coefficient matrix
    # (docs/src/usage/constructors.md:98-116), DualArray bookkeeping (ext/ApproxFunDualNumbersExt.jl:44-63)
    import approxfun_b200 as af
    from oracle import approxfun_oracle as orc
    rng = np.random.default_rng(0)
    for d in (0, 1, 2, 5, 9):
        N = (d + 1) * (d + 2) // 2
        c = rng.standard_normal(N)
        cm = af.api.totaldegree_to_matrix(c)
        assert cm.shape == (d + 1, d + 1)
        assert np.array_equal(cm, orc.tensor_deinterlace(c, d))
        assert np.array_equal(af.api.matrix_to_totaldegree(cm)[:N], c)
        assert np.array_equal(orc.tensor_interlace(cm)[:N], c)
    # a short vector is zero-padded to the next full diagonal: [1,2,3] = 1 + 2 T1(y) + 3 T1(x)
    assert np.array_equal(af.api.totaldegree_to_matrix([1.0, 2.0, 3.0]), [[1.0, 2.0], [3.0, 0.0]])
    dv = af.DualArray([1.0, 2.0], [0.5, 0.25])
    assert dv.shape == (2,) and np.array_equal(af.realpart(dv), [1.0, 2.0]) and np.array_equal(af.dualpart(dv), [0.5, 0.25])
    with pytest.raises(af.DimensionMismatch):
        af.DualArray([1.0, 2.0], [1.0])


def test_library_has_no_undefined_symbols_of_its_own(built):
    # a helper declared in one translation unit and defined in the wrong namespace of another links fine into a shared
    # object and only fails at dlopen on the GPU box: catch it here
    import shutil
    import subprocess

    nm = shutil.which("nm")
    if nm is None:
        pytest.skip("nm not available")
    out = subprocess.run([nm, "-D", "--undefined-only", built._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    own = [ln for ln in out.splitlines() if "afb" in ln]
    assert not own, own
