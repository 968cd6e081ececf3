import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nabladft_b200 import _lib
lib = _lib.load()
M, N, K = (int(x) for x in sys.argv[1:4])
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(3):
    _lib.check(lib.nb200_gemm_tf32x3(M, N, K, _lib.ptr(A), K, _lib.ptr(B), K, 0, _lib.ptr(C), N, 0, None, None, _lib.current_stream()), "g")
torch.cuda.synchronize()
