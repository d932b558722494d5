/* Validation entry points of libppgs_b200.so — NOT part of the product ABI in
 * include/ppgs_b200.h.  They let tests/ exercise one tensor-core kernel in
 * isolation (host buffers in, host buffers out, synchronous). */
#ifndef PPGS_B200_DEBUG_ABI_H_
#define PPGS_B200_DEBUG_ABI_H_
#include "../../include/ppgs_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* out[m][n] = bias[n] + sum_{tap,c} A[m + tap - taps/2][c] * W[n][c][tap]
 * (rows outside [0, M) read as zero), through the tcgen05 GEMM with tile width
 * `bn` (`pair` = 1: the CTA-pair cta_group::2 kernel, needs M % 256 == 0 and bn = 256)
 * and `a_planes` / `b_planes` split-fp16 planes per operand.
 * A: (M, C) fp32, M % 128 == 0, C % 8 == 0.  W: (N, C, taps) fp32 (torch Conv1d
 * layout; (N, C) when taps == 1). */
int ppgs_debug_gemm(ppgs_engine* engine, const float* a_host, const float* w_host,
                    const float* bias_host, int M, int N, int C, int taps, int bn, int pair,
                    int a_planes, int b_planes, float* out_host);

/* One attention call on host data: qkv (rows, 3H) fp32 for one sequence of
 * `tensor_len` rows (rows % 128 == 0) with `valid_len` unmasked keys, `heads` heads of
 * H / heads channels (64, 128 or 256 on the tensor cores). */
int ppgs_debug_attention(ppgs_engine* engine, const float* qkv_host, int rows, int tensor_len,
                         int valid_len, int H, int heads, int causal, int planes,
                         int use_tensor_cores, float* out_host);

/* Copies out and clears the 128 cycle counters the GEMM kernels accumulate when the
 * engine was created with PPGS_B200_TRACE=1 (8 kernel slots x 8 counters, see
 * GemmParams::trace).  Synchronises the device. */
int ppgs_debug_trace(ppgs_engine* engine, unsigned long long* out128);

#ifdef __cplusplus
}
#endif
#endif
