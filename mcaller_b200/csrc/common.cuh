// Shared helpers for libmcaller_b200 (sm_100a).  Device code only sees plain structs of device pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mcaller_b200.h"

void mc_set_error(const char *fmt, ...);

#define MC_CUDA_CHECK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            mc_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return MC_ECUDA;                                                                 \
        }                                                                                    \
    } while (0)

#define MC_LAUNCH_CHECK() MC_CUDA_CHECK(cudaGetLastError())

#define MC_REQUIRE(cond, msg)                       \
    do {                                            \
        if (!(cond)) {                              \
            mc_set_error("%s: %s", __func__, msg);  \
            return MC_EINVAL;                       \
        }                                           \
    } while (0)

// ---- k-mer bit window over a site bitmap: bits [g, g+k) of the bitmap, LSB = position g ---------------
__device__ __forceinline__ uint32_t mc_kmer_bits(const uint32_t *__restrict__ bm, int64_t g, int k) {
    int64_t w = g >> 5;
    int b = (int)(g & 31);
    uint32_t lo = __ldg(bm + w), hi = __ldg(bm + w + 1);
    uint32_t v = __funnelshift_r(lo, hi, b);
    return v & ((1u << k) - 1u);
}

// ---- block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32) -------------
template <int THREADS>
__device__ __forceinline__ int mc_block_exscan(int v, int *s_warp /* [THREADS/32 + 1] */, int &total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int x = (lane < THREADS / 32) ? s_warp[lane] : 0;
        int xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, xi, d);
            if (lane >= d) xi += t;
        }
        if (lane < THREADS / 32) s_warp[lane] = xi - x;
        if (lane == 31) s_warp[THREADS / 32] = xi;
    }
    __syncthreads();
    total = s_warp[THREADS / 32];
    int res = s_warp[wid] + inc - v;
    __syncthreads();
    return res;
}

// generic device exclusive scan (uint32 in -> uint32 out), see scan_util.cu
int mc_exscan_u32(const uint32_t *d_in, uint32_t *d_out, int64_t n, uint64_t *d_total /* may be null */, void *d_ws,
                  cudaStream_t st);
int mc_exscan_u32_dev(const uint32_t *d_in, uint32_t *d_out, int64_t n_cap, const uint64_t *d_n, uint64_t *d_total, void *d_ws,
                      cudaStream_t st);
int64_t mc_exscan_ws_bytes(int64_t n);

// item count kept on the device (stage outputs), clamped to the capacity the launch was sized for
__device__ __forceinline__ int64_t mc_dev_count(const unsigned long long *d_n, int64_t cap) {
    const unsigned long long v = *d_n;
    return v < (unsigned long long)cap ? (int64_t)v : cap;
}

// numpy's pairwise summation (np.add.reduce over a contiguous float64 vector) of f(x_i), i in [i0, i0 + n):
// sequential below 8 values, 8 running lanes + sequential tail up to 128, and above that numpy splits at n/2 rounded down
// to a multiple of 8 and adds the two halves.  The halving is walked with an explicit stack (depth <= log2(n / 64)) instead
// of device-side recursion, whose frames would need more than the default 1 KB of stack for a few thousand values.
template <class F>
__device__ __forceinline__ double mc_pairwise_block(F f, int64_t i0, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, f(i0 + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(i0 + j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(i0 + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(i0 + i));
    return res;
}
template <class F>
__device__ double mc_pairwise_sum(F f, int64_t i0, int64_t n) {
    if (n <= 128) return mc_pairwise_block(f, i0, n);
    constexpr int DEPTH = 40;
    int64_t st_i0[DEPTH], st_n[DEPTH];
    double st_left[DEPTH];
    int st_phase[DEPTH];                      // 0: nothing done, 1: left half running, 2: right half running
    int sp = 0;
    st_i0[0] = i0; st_n[0] = n; st_phase[0] = 0; st_left[0] = 0.0;
    double ret = 0.0;
    for (;;) {
        if (st_phase[sp] == 0) {
            if (st_n[sp] <= 128 || sp + 1 >= DEPTH) {          // leaf (the depth bound cannot bind for n < 2^40)
                ret = mc_pairwise_block(f, st_i0[sp], st_n[sp]);
            } else {
                int64_t n2 = st_n[sp] / 2;
                n2 -= n2 % 8;
                st_phase[sp] = 1;
                st_i0[sp + 1] = st_i0[sp]; st_n[sp + 1] = n2; st_phase[sp + 1] = 0;
                ++sp;
                continue;
            }
        }
        // `ret` holds the sum of the frame at sp: hand it to the parent
        for (;;) {
            if (sp == 0) return ret;
            --sp;
            if (st_phase[sp] == 1) {                             // left half done: run the right half
                int64_t n2 = st_n[sp] / 2;
                n2 -= n2 % 8;
                st_left[sp] = ret;
                st_phase[sp] = 2;
                st_i0[sp + 1] = st_i0[sp] + n2; st_n[sp + 1] = st_n[sp] - n2; st_phase[sp + 1] = 0;
                ++sp;
                break;
            }
            ret = __dadd_rn(st_left[sp], ret);                   // both halves done
        }
    }
}
