"""Tuning aid: per-CTA phase timestamps of tg_solve_kernel (build with -DTG_EXP_TRACE via tools/tune_tridiag.sh and point
OMC_LIB at the variant); writes gpurun_out/tg_trace.npy [n_cta, 16] = start, rng, tile, ascend, scan, lookback, end, smid."""
import ctypes as C
import os
import runpy
import sys

import numpy as np

sys.argv = ["perf_tridiag.py"]
runpy.run_path(os.path.join(os.path.dirname(__file__), "perf_tridiag.py"), run_name="__main__")
from openmcmc_b200 import _cabi

lib = C.CDLL(os.environ["OMC_LIB"])
n = 27840
buf = np.zeros((n, 16), dtype=np.uint64)
rc = lib.omc_debug_trace_read(buf.ctypes.data_as(C.c_void_p), C.c_longlong(buf.nbytes))
assert rc == 0, rc
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/tg_trace.npy", buf)
print("trace saved", buf[:3])
