"""Loads the kernel-source emulator build (tests/emu) -- a developer test tool, never a product path."""
import ctypes
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libafb_emu.so")


def load_binding_module():
    spec = importlib.util.spec_from_file_location("afb_lib_binding", os.path.join(ROOT, "approxfun.jl_b200", "_lib.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["afb_lib_binding"] = mod
    spec.loader.exec_module(mod)
    return mod


def emu_lowlevel():
    so = os.environ.get("AFB_EMU_SO")            # e.g. the AddressSanitizer build (tests/emu/Makefile)
    if so:
        so = os.path.abspath(so)
        subprocess.check_call(["make", "-C", EMU_DIR, os.path.basename(so)], stdout=subprocess.DEVNULL)
    else:
        so = EMU_SO
        subprocess.check_call(["make", "-C", EMU_DIR, "libafb_emu.so"], stdout=subprocess.DEVNULL)
    B = load_binding_module()
    lib = B.bind(ctypes.CDLL(so))
    return B, B.LowLevel(lib)
