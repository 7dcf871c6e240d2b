"""One GEMM shape through wb_debug_gemm_bench (2 warm-up + 3 timed launches): the target of a single-kernel ncu capture, e.g.
  ncu --set full --import-source on -k regex:gemm2_tn_kernel -s 3 -c 1 -o rep python tools/one_gemm.py 96000 2048 512 1     (M N K epilogue)"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from whisper_apr_b200 import _lib
L = _lib.lib()
M, N, K, epi = [int(x) for x in sys.argv[1:5]]
ms = C.c_float(0)
_lib.check(L.wb_debug_gemm_bench(0, 1, M, N, K, epi, 3, C.byref(ms)))
print(ms.value)
