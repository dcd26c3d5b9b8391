"""ctypes loader of tests/csrc/libscgib_probe.so - the tcgen05 hardware probes (test infrastructure; built by
s-cgib_b200/build.py:build_probes, never linked into the product library)."""
import ctypes
import importlib.util
import os
from ctypes import POINTER, c_int, c_int32, c_void_p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "tests", "csrc", "libscgib_probe.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            spec = importlib.util.spec_from_file_location("scgib_build", os.path.join(ROOT, "s-cgib_b200", "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build_probes()
        lib = ctypes.CDLL(PATH)
        lib.scgib_debug_umma.restype = c_int
        lib.scgib_debug_umma.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
        lib.scgib_debug_umma2.restype = c_int
        lib.scgib_debug_umma2.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_void_p]
        lib.scgib_debug_umma_bf16.restype = c_int
        lib.scgib_debug_umma_bf16.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_int32), c_void_p]
        _lib = lib
    return _lib
