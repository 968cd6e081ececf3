// gemm_ps2.cu -- the pre-split-weight GEMM of gemm_ps.cu built with TWO worker groups (tc_pipe.cuh NF_TWO_GROUPS: 8 loader warps run ahead of 8
// epilogue warps): used for K > 128, where every (N tile, K chunk) needs a fresh activation operand and a single group would load it only after
// draining and storing the finished tile.  Same source, second instantiation; exports nb_gemm_ps_impl_2g only.
#define NF_TWO_GROUPS
#define NB_GEMM_PS_2G
#include "gemm_ps.cu"
