// Byte-level parsing helpers shared by the scan kernel (shared-memory bytes) and the record finishing kernel
// (global-memory bytes): contig-name compare, integer / exact-decimal parsing, NNNNNN test.
#pragma once
#include "common.cuh"

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// ---- byte sources: warp-private shared memory (fast path) or global memory (slow path) ---------------------------------
struct SmemBytes {
    const uint8_t *p;
    __device__ __forceinline__ int operator[](int i) const { return p[i]; }
};
struct GlobalBytes {
    const uint8_t *p;
    int64_t limit;     // bytes readable from p
    __device__ __forceinline__ int operator[](int64_t i) const { return i < limit ? __ldg(p + i) : 0x0a; }
};

template <class B>
__device__ __forceinline__ bool contig_match(const B &t, int q, const mc_refindex &R, int cid) {
    const int o0 = __ldg(R.d_name_off + cid), L = __ldg(R.d_name_off + cid + 1) - o0;
    for (int j = 0; j < L; ++j)
        if (t[q + j] != __ldg(R.d_names + o0 + j)) return false;
    return t[q + L] <= 0x20;
}
template <class B>
__device__ __forceinline__ int contig_cmp(const B &t, int q, const mc_refindex &R, int cid) {
    const int o0 = __ldg(R.d_name_off + cid), L = __ldg(R.d_name_off + cid + 1) - o0;
    for (int j = 0; j < L; ++j) {
        const int a = t[q + j], b = __ldg(R.d_names + o0 + j);
        if (a <= 0x20) return -1;
        if (a != b) return a - b;
    }
    return t[q + L] <= 0x20 ? 0 : 1;
}
// contigs are sorted by name on the host: hint first, then binary search
template <class B>
__device__ __forceinline__ int contig_lookup(const B &t, int q, const mc_refindex &R, int hint) {
    if (contig_match(t, q, R, hint)) return hint;
    int lo = 0, hi = R.n_contigs - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int c = contig_cmp(t, q, R, mid);
        if (c == 0) return mid;
        if (c < 0) hi = mid - 1; else lo = mid + 1;
    }
    return -1;
}
template <class B>
__device__ __forceinline__ bool parse_uint(const B &t, int q, int &out) {
    int v = 0, nd = 0, c;
    while ((c = t[q]) >= '0' && c <= '9') { v = v * 10 + (c - '0'); ++nd; ++q; if (nd > 9) return false; }
    if (nd == 0 || c > 0x20) return false;
    out = v;
    return true;
}
template <class B>
__device__ __forceinline__ bool parse_int(const B &t, int q, int &out) {
    bool neg = false;
    if (t[q] == '-') { neg = true; ++q; } else if (t[q] == '+') ++q;
    int v;
    if (!parse_uint(t, q, v)) return false;
    out = neg ? -v : v;
    return true;
}
// plain decimal -> correctly rounded double (mantissa <= 2^53, <= 18 digits: one exact division)
template <class B>
__device__ __forceinline__ bool parse_decimal(const B &t, int q, double &out) {
    bool neg = false;
    int c = t[q];
    if (c == '-') { neg = true; ++q; } else if (c == '+') ++q;
    unsigned long long m = 0;
    int nd = 0, nfrac = 0;
    while ((c = t[q]) >= '0' && c <= '9') { m = m * 10ull + (unsigned)(c - '0'); ++nd; ++q; if (nd > 18) return false; }
    if (c == '.') {
        ++q;
        while ((c = t[q]) >= '0' && c <= '9') { m = m * 10ull + (unsigned)(c - '0'); ++nd; ++nfrac; ++q; if (nd > 18) return false; }
    }
    if (nd == 0 || c > 0x20 || m > (1ull << 53)) return false;
    const double v = __ddiv_rn((double)m, c_pow10[nfrac]);
    out = neg ? -v : v;
    return true;
}
template <class B>
__device__ __forceinline__ bool is_nnnnnn(const B &t, int q) {
    if (t[q] != 'N') return false;
    return t[q + 1] == 'N' && t[q + 2] == 'N' && t[q + 3] == 'N' && t[q + 4] == 'N' && t[q + 5] == 'N' && t[q + 6] <= 0x20;
}


// reference_kmer (col 3) == model_kmer (col 10): token compare up to the first whitespace on either side
template <class B>
__device__ __forceinline__ bool tokens_equal(const B &t, int64_t a, int64_t b) {
    for (;;) {
        const int ca = t[a++], cb = t[b++];
        const bool ea = ca <= 0x20, eb = cb <= 0x20;
        if (ea || eb) return ea && eb;
        if (ca != cb) return false;
    }
}

// event index, np.round(event_mean - model_mean, 4) and the k-mer equality flag of one line (extract_contexts.py:150,
// :169, :286); f5/f6/f2/f9/f10 are byte offsets of the columns from the start of `t`
template <class B>
__device__ __forceinline__ void parse_values(const B &t, int f2, int f5, int f6, int f9, int f10, int &event_idx, double &diff,
                                             uint32_t &flags) {
    event_idx = 0;
    if (!parse_int(t, f5, event_idx)) flags |= MC_RF_BADIDX;
    double ev = 0.0, md = 0.0;
    if (!parse_decimal(t, f6, ev) || !parse_decimal(t, f10, md)) { flags |= MC_RF_BADNUM; diff = 0.0; }
    else diff = __ddiv_rn(rint(__dmul_rn(__dsub_rn(ev, md), 1e4)), 1e4);
    if (tokens_equal(t, f2, f9)) flags |= MC_RF_EQ;
}

// ---- 8-byte register fast paths for the values of a record (any other shape falls back to the byte loops of parse.cuh) ----
// 0x80 per byte of v that is not an ASCII digit
__device__ __forceinline__ unsigned long long nondigit8(unsigned long long v) {
    const unsigned long long x = v ^ 0x3030303030303030ull;
    return (((x & 0x7f7f7f7f7f7f7f7full) + 0x7676767676767676ull) | v) & 0x8080808080808080ull;
}
// n (1..7) digit bytes at the low end of v -> value
__device__ __forceinline__ uint32_t digits_value(unsigned long long v, int n) {
    const unsigned long long x = (v ^ 0x3030303030303030ull) << (64 - 8 * n);      // right-aligned digit values, zeros below
    const uint32_t L = (uint32_t)x, H = (uint32_t)(x >> 32);
    auto conv4 = [](uint32_t h) {                                 // 4 digit values, most significant in byte 0 -> 0..9999
        const uint32_t t = ((h * 2561u) >> 8) & 0x00ff00ffu;
        return (t * 6553601u) >> 16;
    };
    return conv4(L) * 10000u + conv4(H);
}
// "<1..7 digits><ws>" -> value; false for any other shape
__device__ __forceinline__ bool fast_uint8(unsigned long long v, int &out) {
    const unsigned long long nd = nondigit8(v);
    if (nd == 0ull) return false;
    const int n = (__ffsll((long long)nd) - 1) >> 3;
    if (n == 0 || ((v >> (8 * n)) & 0xFFull) > 0x20ull) return false;
    out = (int)digits_value(v, n);
    return true;
}
// "<digits>.<digits><ws>" with at most 7 bytes before the whitespace -> mantissa and number of fraction digits
__device__ __forceinline__ bool fast_decimal8(unsigned long long v, uint32_t &mant, int &nfrac) {
    const unsigned long long nd = nondigit8(v);
    if (nd == 0ull) return false;
    const int p1 = (__ffsll((long long)nd) - 1) >> 3;             // first non-digit: must be the point
    if (p1 == 0 || p1 > 6 || ((v >> (8 * p1)) & 0xFFull) != 0x2eull) return false;
    const unsigned long long low = (1ull << (8 * p1)) - 1ull;
    const unsigned long long w = (v & low) | ((v >> 8) & ~low);    // the point removed: 7 bytes
    const unsigned long long nd2 = ((nd & low) | ((nd >> 8) & ~low)) & 0x0080808080808080ull;
    if (nd2 == 0ull) return false;
    const int p2 = (__ffsll((long long)nd2) - 1) >> 3;            // first non-digit after it: must be whitespace
    if (((w >> (8 * p2)) & 0xFFull) > 0x20ull) return false;
    mant = digits_value(w, p2);
    nfrac = p2 - p1;
    return true;
}
// token at pa == token at pb for tokens of up to 7 bytes: 1 / 0, or -1 when one of them is longer (byte loop decides)
__device__ __forceinline__ int fast_tokens_equal8(unsigned long long a, unsigned long long b) {
    auto ws8 = [](unsigned long long v) {                          // 0x80 per byte <= 0x20
        return ~((((v & 0x7f7f7f7f7f7f7f7full) + 0x5f5f5f5f5f5f5f5full) | v)) & 0x8080808080808080ull;
    };
    const unsigned long long wa = ws8(a), wb = ws8(b);
    if (wa == 0ull || wb == 0ull) return -1;
    const int la = (__ffsll((long long)wa) - 1) >> 3, lb = (__ffsll((long long)wb) - 1) >> 3;
    if (la != lb) return 0;
    return (((a ^ b) & ((1ull << (8 * la)) - 1ull)) == 0ull) ? 1 : 0;
}
