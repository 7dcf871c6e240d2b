import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
from whisper_apr_b200 import _lib
L = _lib.lib()
M, N, K, epi = [int(x) for x in sys.argv[1:5]]
ms = C.c_float(0)
_lib.check(L.wb_debug_gemm_bench(0, 1, M, N, K, epi, 3, C.byref(ms)))
print(ms.value)
