"""Python handle on the CPU oracle (oracle/librem2d_oracle.so). TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs —
never by the gym_rem2d_b200 package. It reuses the library-agnostic ctypes binding because the oracle
exports the same C-ABI as the CUDA library (include/rem2d.h).
"""
import ctypes as C
import os
import subprocess

from gym_rem2d_b200.capi import Engine, load_library

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "librem2d_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "rem2d_oracle.c")
    if force or not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B" if force else "-s"])
    return ORACLE_LIB


class OracleEngine(Engine):
    def __init__(self, threads=1, **overrides):
        super().__init__(lib_path=build(), **overrides)
        self.lib.rem2d_oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        self.lib.rem2d_oracle_set_threads(self.h, int(threads))
