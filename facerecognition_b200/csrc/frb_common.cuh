// frb_common.cuh — shared helpers for the libfrb200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>

#include "../../include/frb200.h"

namespace frb {

// thread-local last-error text behind frb_last_error()
char *last_error_buf();
void set_error(const char *fmt, ...);

int sm_count();  // cached multiProcessorCount of the current device

// frb_normalize_rows with optional per-row scratch arrays that the same launch resets (-inf floats, zero ints)
int normalize_rows_impl(const float *x, int64_t rows, int dim, int mode, void *out, int out_dtype, float *neg_inf_fill,
                        int *zero_fill, cudaStream_t st);

// merge of compact per-query candidate buffers: cand[q * cap + i], i < cnt[q]  (core.cu)
int topk_merge_compact(const float *cs, const int64_t *ci, const int *cnt, int64_t cap, int64_t n_query, int k, int largest,
                       float *os, int64_t *oi, cudaStream_t st);

#define FRB_CHECK_ARG(cond, ...)                \
    do {                                        \
        if (!(cond)) {                          \
            ::frb::set_error(__VA_ARGS__);      \
            return FRB_ERR_INVALID;             \
        }                                       \
    } while (0)

#define FRB_CUDA_OK(expr)                                                                            \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            ::frb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FRB_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

#define FRB_LAUNCH_OK(name)                                                                    \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) {                                                               \
            ::frb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));         \
            return FRB_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// Event bracket around one launch of a hot kernel (no-op unless frb_profile_enable(1)).
struct ProfileScope {
    int kernel;
    cudaStream_t stream;
    cudaEvent_t start, stop;
    bool on;
    ProfileScope(int kernel_id, cudaStream_t st);
    ~ProfileScope();
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- candidate lists -----------------------------------------------------------------------
// A running best-k list for one query, filled from a stream whose row index only increases.
// kth is the admission threshold kept in a register; the list itself lives in local/shared
// memory and is touched only on the (rare) insertions.  LARGEST: descending scores (cosine);
// else ascending distances (chi-square).  Equal keys keep the earlier (lower) row first.
template <bool LARGEST>
__device__ __forceinline__ bool beats(float v, float kth)
{
    return LARGEST ? (v > kth) : (v < kth);
}

template <bool LARGEST>
__device__ __forceinline__ float worst_value()
{
    return LARGEST ? -INFINITY : INFINITY;
}

template <bool LARGEST>
__device__ __forceinline__ void list_init(float *s, int64_t *id, int k)
{
    for (int i = 0; i < k; i++) {
        s[i] = worst_value<LARGEST>();
        id[i] = -1;
    }
}

// Insert (v, idx); caller guarantees beats(v, s[k-1]) and that idx is larger than every idx
// already in the list with an equal key (streaming order).  Returns the new threshold.
template <bool LARGEST>
__device__ __forceinline__ float list_insert_stream(float *s, int64_t *id, int k, float v, int64_t idx)
{
    int p = k - 1;
    while (p > 0 && beats<LARGEST>(v, s[p - 1])) {
        s[p] = s[p - 1];
        id[p] = id[p - 1];
        --p;
    }
    s[p] = v;
    id[p] = idx;
    return s[k - 1];
}

// Total order for merging lists that arrive in arbitrary order: key first, then lower idx;
// padding (idx < 0) is worse than everything.
template <bool LARGEST>
__device__ __forceinline__ bool better(float v, int64_t idx, float w, int64_t jdx)
{
    if (idx < 0) return false;
    if (jdx < 0) return true;
    if (v != w) return LARGEST ? (v > w) : (v < w);
    return idx < jdx;
}

}  // namespace frb
