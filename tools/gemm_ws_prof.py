"""Role timing of the warp-specialised tcgen05 GEMM (build with NB200_NVCC_EXTRA=-DNB_WS_PROF).
Prints, per CTA and per N tile, the cycles each role spent waiting: tells which of producer / issuer / epilogue paces the kernel."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nabladft_b200 import _lib
lib = _lib.load()
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (100096, 8320, 128)
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
fn = lib.nb200_debug_ws_prof; fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]; fn.restype = ctypes.c_int
buf = (ctypes.c_ulonglong * 8)()
def run():
    _lib.check(lib.nb200_gemm_tf32x3(M, N, K, _lib.ptr(A), K, _lib.ptr(B), K, 0, _lib.ptr(C), N, 0, None, None, _lib.current_stream()), "g")
run(); torch.cuda.synchronize(); fn(buf, 1)
run(); torch.cuda.synchronize(); fn(buf, 1)
v = list(buf); ctas = max(v[7], 1); tiles = (N + 127) // 128
names = ["producer wait empty", "producer total", "issuer wait full", "issuer wait acc_empty", "issuer total", "epilogue wait acc_full", "epilogue total"]
for n, x in zip(names, v):
    print(f"{n:26s} {x / ctas / tiles:9.1f} cycles / tile")
print("CTAs", ctas, "tiles/CTA", tiles)
