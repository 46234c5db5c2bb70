// Stage 1 (K1): eventalign TSV tokeniser + line filter.  One WARP per 3712-byte text chunk, no block barriers.
//
// Replaces the reader / tokeniser / per-line filters of the reference's extract_features
// (extract_contexts.py:140-176): readlines + line.split()[:12], the '<12 fields' drop (:149-152),
// the contig lookup (:154-160), the NNNNNN drop (:167) and the "does this k-mer touch an 'M'"
// test that gates everything after (:176, :242, :269).
//
// Per chunk (persistent warps stride over the chunks of the text):
//   1. each lane pulls 4 x 32 B of the chunk (+32 B look-behind, +352 B look-ahead) with 16 B loads, classifies
//      the bytes in registers (SWAR compare + IDP.4A bit packing) into two bit maps -- non-whitespace (byte > 0x20)
//      and newline (byte == 0x0a) -- and parks text and bit maps in the warp's private shared-memory slice;
//   2. field starts = nonws & ~(nonws << 1); line starts inside the chunk are compacted into a list with one warp scan;
//   3. one lane per line: popcount-select on the field-start bits finds columns 2, 10 and 12 without touching the
//      bytes in between (the ~58 B read name is never walked), the contig is resolved against a warp-uniform hint,
//      column 2 is parsed and the per-position candidate bitmap (L1/L2 resident) is tested;
//   4. warp ballots decide which lines matter (candidate, successor of a candidate, first kept line of the chunk, or
//      every kept line in dense mode); only those are fully parsed (event index, currents as exact decimals ->
//      float64 diff rounded like np.round(x, 4), k-mer equality, read-name span) into 32-byte records.  Record slots
//      are reserved in blocks of 64 per warp, so the global allocation counter sees ~1 atomic per 20 chunks.
// Lines whose first 12 columns do not fit the look-ahead take a byte-wise slow path straight from global memory.
// Algorithmic HBM traffic: the text itself (once) + 32 B per record (~3 B per line in sparse mode).
#include "common.cuh"

namespace {

constexpr int CHUNK = MC_TILE_BYTES;          // 3712 = 29 * 128: ~29 lines of ~128 B, one per lane
constexpr int LOOKB = 32;
constexpr int LOOKA = 352;
constexpr int WB = LOOKB + CHUNK + LOOKA;     // 4096 bytes staged per chunk
constexpr int NW = WB / 32;                   // 128 mask words = 4 per lane
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int LCAP = 64;                      // line-list capacity per pass
constexpr int ECAP = 192;                     // queued record lines per chunk (a chunk holds <= 3712/23 keepable lines)
#ifndef MC_SCAN_MIN_CTAS
#define MC_SCAN_MIN_CTAS 3
#endif
constexpr int RESERVE = 256;                   // record slots reserved per global atomic
static_assert(WB == 4096 && NW == 128, "chunk geometry");
static_assert(CHUNK % 16 == 0, "chunks must keep 16-byte alignment");
static_assert(MC_TEXT_PAD >= LOOKA + 64, "text padding must cover the look-ahead");

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

struct WarpSmem {
    alignas(16) uint8_t text[WB + 16];
    uint32_t nonws[NW + 4];
    uint32_t fs[NW + 4];
    uint32_t nl[NW + 4];
    uint32_t ls[NW + 4];
    uint16_t lstart[LCAP + 2];
};

// ---- byte classification ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gt20_msb(uint32_t w) { return (((w & 0x7f7f7f7fu) + 0x5f5f5f5fu) | w) & 0x80808080u; }
__device__ __forceinline__ uint32_t eq0a_msb(uint32_t w) {
    const uint32_t x = w ^ 0x0a0a0a0au;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// 8 msb-form words (0x80 per flagged byte) -> 32 flag bits in byte order.  IDP.4A sums 0x80 * weight per byte: two
// words fill bits 7..14 of one accumulator.
__device__ __forceinline__ uint32_t pack32(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3, uint32_t m4, uint32_t m5,
                                           uint32_t m6, uint32_t m7) {
    const uint32_t lo = 0x08040201u, hi = 0x80402010u;
    const uint32_t a = __dp4a(m1, hi, __dp4a(m0, lo, 0u));
    const uint32_t b = __dp4a(m3, hi, __dp4a(m2, lo, 0u));
    const uint32_t c = __dp4a(m5, hi, __dp4a(m4, lo, 0u));
    const uint32_t d = __dp4a(m7, hi, __dp4a(m6, lo, 0u));
    return (a >> 7) + b * 2u + c * 512u + d * 131072u;
}

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

// ---- field location on the field-start bit map --------------------------------------------------------------------------
struct FsCursor {
    const uint32_t *fs;
    int w;          // current word
    uint32_t m;     // unconsumed bits of the current word
    int before;     // field starts consumed in earlier words
    __device__ __forceinline__ void init(const uint32_t *fs_, int s) { fs = fs_; w = s >> 5; m = fs[w] & (0xFFFFFFFFu << (s & 31)); before = 0; }
    // position of the n-th (0-based, counted from the line start) field start; n must not decrease.  NW*32 if it lies
    // beyond the staged bytes.
    __device__ __forceinline__ int find(int n) {
        int c = __popc(m);
        while (before + c <= n) {
            before += c;
            if (++w >= NW) { m = 0u; return NW * 32; }
            m = fs[w];
            c = __popc(m);
        }
        uint32_t t = m;
        for (int j = n - before; j > 0; --j) t &= t - 1u;
        return (w << 5) + __ffs(t) - 1;
    }
};
__device__ __forceinline__ int next_bit(const uint32_t *m, int q) {
    int w = q >> 5;
    uint32_t v = m[w] & (0xFFFFFFFFu << (q & 31));
    while (v == 0u) {
        if (++w >= NW) return NW * 32;
        v = m[w];
    }
    return (w << 5) + __ffs(v) - 1;
}
__device__ __forceinline__ int token_end(const uint32_t *nonws, int q) {
    int w = q >> 5;
    uint32_t v = ~nonws[w] & (0xFFFFFFFFu << (q & 31));
    while (v == 0u) {
        if (++w >= NW) return NW * 32;
        v = ~nonws[w];
    }
    return (w << 5) + __ffs(v) - 1;
}

// result of the structural parse of one line
struct LineHead {
    int cid, pos;
    uint32_t status;     // ST_*
};
enum { ST_KEPT = 1u, ST_CAND = 2u, ST_SHORT = 4u, ST_UNKNOWN = 8u, ST_NNN = 16u, ST_BADPOS = 32u, ST_SLOW = 64u };

template <class B>
__device__ __forceinline__ void classify_line(const B &t, int f0, int f1, int f9, const mc_refindex &R, int hint, int64_t hint_base,
                                              int hint_len, LineHead &L) {
    L.cid = contig_lookup(t, f0, R, hint);
    if (L.cid < 0) { L.status = ST_UNKNOWN; return; }
    if (is_nnnnnn(t, f9)) { L.status = ST_NNN; return; }
    if (!parse_uint(t, f1, L.pos)) { L.status = ST_BADPOS; return; }
    L.status = ST_KEPT;
    const int len = (L.cid == hint) ? hint_len : __ldg(R.d_len + L.cid);
    if (L.pos < len) {
        const int64_t g = ((L.cid == hint) ? hint_base : __ldg(R.d_base + L.cid)) + L.pos;
        if ((__ldg(R.d_cand + (g >> 5)) >> (g & 31)) & 1u) L.status |= ST_CAND;
    }
}

// full parse of an emitted line into a record; F[] = starts of fields 2,3,5,6,9,10, name_end = end of field 3
template <class B>
__device__ __forceinline__ void fill_record(const B &t, int64_t line_goff, int s, int f2, int f3, int name_end, int f5, int f6, int f9,
                                            int f10, const LineHead &L, mc_record &r) {
    r.line_lo = (uint32_t)(line_goff & 0xFFFFFFFFll);
    r.line_hi = (uint16_t)(line_goff >> 32);
    r.name_off = (uint16_t)(f3 - s);
    r.name_len = (uint16_t)(name_end - f3);
    r.pos = L.pos;
    r.contig = (uint16_t)L.cid;
    uint32_t fl = (L.status & ST_CAND) ? MC_RF_CAND : 0u;
    int ev_idx = 0;
    if (!parse_int(t, f5, ev_idx)) fl |= MC_RF_BADIDX;
    r.event_idx = ev_idx;
    double ev = 0.0, md = 0.0;
    if (!parse_decimal(t, f6, ev) || !parse_decimal(t, f10, md)) { fl |= MC_RF_BADNUM; r.diff = 0.0; }
    else r.diff = __ddiv_rn(rint(__dmul_rn(__dsub_rn(ev, md), 1e4)), 1e4);          // np.round(ev - model, 4)
    {   // reference_kmer (col 3) == model_kmer (col 10)
        int a = f2, b = f9;
        bool eq = true;
        for (;;) {
            const int ca = t[a++], cb = t[b++];
            const bool ea = ca <= 0x20, eb = cb <= 0x20;
            if (ea || eb) { eq = ea && eb; break; }
            if (ca != cb) { eq = false; break; }
        }
        if (eq) fl |= MC_RF_EQ;
    }
    r.flags = (uint8_t)fl;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
}

// byte-wise field finder for the slow path: starts of fields 0..11 of the line at t[0..), -1 when the line ends first
__device__ __noinline__ int slow_fields(const GlobalBytes &t, int64_t *f, int64_t *name_end) {
    int nf = 0;
    bool in_tok = false;
    for (int64_t i = 0;; ++i) {
        const int c = t[i];
        if (c == 0x0a) { if (in_tok && nf == 4) *name_end = i; break; }
        const bool ws = c <= 0x20;
        if (!ws && !in_tok) { if (nf < 12) f[nf] = i; ++nf; in_tok = true; if (nf >= 13) break; }
        else if (ws && in_tok) { in_tok = false; if (nf == 4) *name_end = i; if (nf >= 12) break; }
    }
    return nf;
}

// starts of the first 12 fields of the line at smem offset s, from the field-start bit map (NW*32 where the staged
// bytes end first)
__device__ __forceinline__ void extract_fields(const uint32_t *fs, int s, int (&F)[12]) {
    int w = s >> 5;
    uint32_t m = fs[w] & (0xFFFFFFFFu << (s & 31));
    bool dead = false;
#pragma unroll
    for (int f = 0; f < 12; ++f) {
        while (m == 0u && !dead) {
            if (++w >= NW) dead = true; else m = fs[w];
        }
        if (dead) F[f] = NW * 32;
        else { F[f] = (w << 5) + __ffs(m) - 1; m &= m - 1u; }
    }
}

// first 8 bytes at smem offset q (any alignment) as a little-endian u64
__device__ __forceinline__ unsigned long long load8(const uint8_t *text, int q) {
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(text + (q & ~3));
    const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
    const int sh = 8 * (q & 3);
    const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
    return ((unsigned long long)hi << 32) | lo;
}

// the whole line handled from global memory (first 12 columns outran the staged look-ahead): status + record
__device__ __noinline__ void line_from_global(const uint8_t *d_text, int64_t nbytes, int64_t goff, const mc_refindex &R, int hint,
                                              int64_t hint_base, int hint_len, LineHead &L, mc_record &rec) {
    const GlobalBytes GB{d_text + goff, nbytes + MC_TEXT_PAD - 64 - goff};
    int64_t sf[12], name_end = 0;
    const int nf = slow_fields(GB, sf, &name_end);
    if (nf < 12) { L.status = ST_SHORT; return; }
    classify_line(GB, (int)sf[0], (int)sf[1], (int)sf[9], R, hint, hint_base, hint_len, L);
    if (L.status & ST_KEPT)
        fill_record(GB, goff, 0, (int)sf[2], (int)sf[3], (int)name_end, (int)sf[5], (int)sf[6], (int)sf[9], (int)sf[10], L, rec);
    L.status |= ST_SLOW;
}

__global__ void __launch_bounds__(THREADS, MC_SCAN_MIN_CTAS)
k_scan(const uint8_t *__restrict__ d_text, int64_t nbytes, int64_t text_limit16, int64_t n_chunks, mc_refindex R, int dense,
       mc_record *__restrict__ d_rec, unsigned long long rec_cap, uint32_t *__restrict__ d_tile_tab,
       unsigned long long *__restrict__ d_counters) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    WarpSmem &S = reinterpret_cast<WarpSmem *>(smem_raw)[wib];
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + wib;
    const int64_t warp_stride = (int64_t)gridDim.x * WARPS;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // warp-uniform state
    int hint = -1;
    int64_t hint_base = 0;
    int hint_len = 0, hint_nlen = 0;
    unsigned long long hint_key = 0ull, hint_keymask = 0ull;      // first bytes of the hint contig's name (fast compare)
    auto set_hint = [&](int c) {
        hint = c;
        hint_base = __ldg(R.d_base + c);
        hint_len = __ldg(R.d_len + c);
        const int o0 = __ldg(R.d_name_off + c);
        hint_nlen = __ldg(R.d_name_off + c + 1) - o0;
        hint_key = 0ull;
        for (int j = 0; j < hint_nlen && j < 7; ++j) hint_key |= (unsigned long long)__ldg(R.d_names + o0 + j) << (8 * j);
        hint_keymask = hint_nlen <= 7 ? ((1ull << (8 * hint_nlen)) - 1ull) : 0ull;
    };
    set_hint(0);
    unsigned long long slot_cur = 0ull, slot_end = 0ull;          // reserved record slots [cur, end)
    // per-lane counters, reduced once at the end
    unsigned c_lines = 0, c_kept = 0, c_short = 0, c_unknown = 0, c_nnn = 0, c_badpos = 0, c_slow = 0, c_overflow = 0;

    const SmemBytes T{S.text};

    for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warp_stride) {
        const int64_t G0 = chunk * (int64_t)CHUNK - LOOKB;        // global offset of smem byte 0
        const bool interior = G0 >= 0 && G0 + WB <= text_limit16 && G0 + WB <= nbytes;
        __syncwarp();
        // ---- 1. load + classify: lane owns words lane, lane+32, lane+64, lane+96 -----------------------------------
        uint32_t prev_nonws_top = 0u, prev_nl_top = 0u;           // top bits of word 32r-1 (lane 31 of the previous round)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint4 va[2], vb[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = 2 * half + rr;
                const int64_t g = G0 + 32 * (32 * r + lane);
                if (interior) {
                    va[rr] = __ldcs(reinterpret_cast<const uint4 *>(d_text + g));
                    vb[rr] = __ldcs(reinterpret_cast<const uint4 *>(d_text + g + 16));
                } else {
                    const uint4 nlv = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
                    va[rr] = (g >= 0 && g < text_limit16) ? __ldcs(reinterpret_cast<const uint4 *>(d_text + g)) : nlv;
                    vb[rr] = (g + 16 >= 0 && g + 16 < text_limit16) ? __ldcs(reinterpret_cast<const uint4 *>(d_text + g + 16)) : nlv;
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = 2 * half + rr;
                const int w = 32 * r + lane;
                *reinterpret_cast<uint4 *>(S.text + 32 * w) = va[rr];
                *reinterpret_cast<uint4 *>(S.text + 32 * w + 16) = vb[rr];
                const uint32_t nonws = pack32(gt20_msb(va[rr].x), gt20_msb(va[rr].y), gt20_msb(va[rr].z), gt20_msb(va[rr].w),
                                              gt20_msb(vb[rr].x), gt20_msb(vb[rr].y), gt20_msb(vb[rr].z), gt20_msb(vb[rr].w));
                const uint32_t nl = pack32(eq0a_msb(va[rr].x), eq0a_msb(va[rr].y), eq0a_msb(va[rr].z), eq0a_msb(va[rr].w),
                                           eq0a_msb(vb[rr].x), eq0a_msb(vb[rr].y), eq0a_msb(vb[rr].z), eq0a_msb(vb[rr].w));
                // top bits of the previous word: lane-1 of this round, or lane 31 of the previous round for lane 0
                uint32_t pn = __shfl_up_sync(0xffffffffu, nonws >> 31, 1);
                uint32_t pl = __shfl_up_sync(0xffffffffu, nl >> 31, 1);
                if (lane == 0) { pn = prev_nonws_top; pl = prev_nl_top; }
                prev_nonws_top = __shfl_sync(0xffffffffu, nonws >> 31, 31);
                prev_nl_top = __shfl_sync(0xffffffffu, nl >> 31, 31);
                S.nonws[w] = nonws;
                S.nl[w] = nl;
                S.fs[w] = nonws & ~((nonws << 1) | pn);
                // line starts: byte p starts a line iff byte p-1 is '\n'; owned range LOOKB <= p < LOOKB+CHUNK, global p < nbytes
                uint32_t ls = (nl << 1) | pl;
                const int p0 = 32 * w;
                if (p0 < LOOKB || p0 >= LOOKB + CHUNK) ls = 0u;   // LOOKB and CHUNK are multiples of 32: whole words
                else if (!interior) {
                    const int64_t room = nbytes - (G0 + p0);
                    if (room <= 0) ls = 0u;
                    else if (room < 32) ls &= (1u << room) - 1u;
                }
                S.ls[w] = ls;
            }
        }
        if (lane == 0) { S.nonws[NW] = 0u; S.fs[NW] = 0u; S.nl[NW] = 0xFFFFFFFFu; }
        __syncwarp();

        // ---- 2. ordered line list: lane owns words 4*lane .. 4*lane+3 -------------------------------------------------
        const uint4 lsv = *reinterpret_cast<const uint4 *>(&S.ls[4 * lane]);
        const int my_cnt = __popc(lsv.x) + __popc(lsv.y) + __popc(lsv.z) + __popc(lsv.w);
        int incl = my_cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int tt = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += tt;
        }
        const int total_lines = __shfl_sync(0xffffffffu, incl, 31);
        const int my_first = incl - my_cnt;
        c_lines += (lane == 0) ? (unsigned)total_lines : 0u;

        // record slots for this chunk are taken from the warp's reserved block; make sure it can hold every line
        if (slot_end - slot_cur < (unsigned long long)total_lines) {
            unsigned long long got = 0ull;
            const unsigned want = total_lines > RESERVE ? (unsigned)total_lines : (unsigned)RESERVE;
            if (lane == 0) got = atomicAdd(&d_counters[MC_C_RECORDS], (unsigned long long)want);
            slot_cur = __shfl_sync(0xffffffffu, got, 0);
            slot_end = slot_cur + want;
        }
        const unsigned long long chunk_base = slot_cur;

        int prev_state = -1;       // -1: no kept line yet in this chunk, 0: last kept line not a candidate, 1: candidate
        for (int pass0 = 0; pass0 < total_lines; pass0 += 32) {
            // list entries [pass0, pass0+33) (one extra so every lane knows where its line ends)
            {
                int idx = my_first - pass0;
                const uint32_t wv[4] = {lsv.x, lsv.y, lsv.z, lsv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t m = wv[j];
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1u;
                        if (idx >= 0 && idx <= 32) S.lstart[idx] = (uint16_t)(32 * (4 * lane + j) + b);
                        ++idx;
                    }
                }
            }
            __syncwarp();
            const int n_pass = min(32, total_lines - pass0);
            // ---- 3. structural parse, one lane per line --------------------------------------------------------------
            LineHead L;
            L.cid = -1; L.pos = 0; L.status = 0u;
            int s = 0;
            int F[12];
            alignas(16) mc_record rec;
            if (lane < n_pass) {
                s = S.lstart[lane];
                const int e = (pass0 + lane + 1 < total_lines) ? (int)S.lstart[lane + 1] - 1 : next_bit(S.nl, s);
                extract_fields(S.fs, s, F);
                if (F[11] < e) {
                    // contig: compare the first bytes with the warp's hint in registers, full lookup on a miss
                    const unsigned long long k8 = load8(S.text, F[0]);
                    if (hint_keymask && (k8 & hint_keymask) == hint_key && ((k8 >> (8 * hint_nlen)) & 0xFFull) <= 0x20ull) L.cid = hint;
                    else L.cid = contig_lookup(T, F[0], R, hint);
                    if (L.cid < 0) L.status = ST_UNKNOWN;
                    else if (is_nnnnnn(T, F[9])) L.status = ST_NNN;
                    else if (!parse_uint(T, F[1], L.pos)) L.status = ST_BADPOS;
                    else {
                        L.status = ST_KEPT;
                        const int len = (L.cid == hint) ? hint_len : __ldg(R.d_len + L.cid);
                        if (L.pos < len) {
                            const int64_t g = ((L.cid == hint) ? hint_base : __ldg(R.d_base + L.cid)) + L.pos;
                            if ((__ldg(R.d_cand + (g >> 5)) >> (g & 31)) & 1u) L.status |= ST_CAND;
                        }
                    }
                } else if (e >= NW * 32 || F[11] >= NW * 32) {
                    line_from_global(d_text, nbytes, G0 + s, R, hint, hint_base, hint_len, L, rec);
                    ++c_slow;
                } else {
                    L.status = ST_SHORT;
                }
            }
            c_kept += (L.status & ST_KEPT) ? 1u : 0u;
            c_short += (L.status & ST_SHORT) ? 1u : 0u;
            c_unknown += (L.status & ST_UNKNOWN) ? 1u : 0u;
            c_nnn += (L.status & ST_NNN) ? 1u : 0u;
            c_badpos += (L.status & ST_BADPOS) ? 1u : 0u;

            // ---- 4. which lines matter: ballots over the 32 lines of this pass -----------------------------------------
            const uint32_t kept_m = __ballot_sync(0xffffffffu, (L.status & ST_KEPT) != 0u);
            const uint32_t cand_m = __ballot_sync(0xffffffffu, (L.status & ST_CAND) != 0u);
            bool emit = false;
            if (L.status & ST_KEPT) {
                if (dense || (L.status & ST_CAND)) emit = true;
                else {
                    const uint32_t below = kept_m & lt_mask;
                    const int st = below ? (int)((cand_m >> (31 - __clz(below))) & 1u) : prev_state;
                    emit = (st != 0);            // predecessor is a candidate, or this is the first kept line of the chunk
                }
            }
            const uint32_t emit_m = __ballot_sync(0xffffffffu, emit);
            if (kept_m) {
                const int top = 31 - __clz(kept_m);
                prev_state = (int)((cand_m >> top) & 1u);
                const int new_hint = __shfl_sync(0xffffffffu, L.cid, top);     // contig hint follows the last kept line
                if (new_hint != hint) set_hint(new_hint);
            }
            // ---- 5. records: the emitting lanes finish the parse (values) and store ---------------------------------------
            if (emit) {
                if (!(L.status & ST_SLOW)) fill_record(T, G0 + s, s, F[2], F[3], token_end(S.nonws, F[3]), F[5], F[6], F[9], F[10], L, rec);
                const unsigned long long slot = slot_cur + __popc(emit_m & lt_mask);
                if (slot < rec_cap) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(&rec);
                    uint4 *dst = reinterpret_cast<uint4 *>(d_rec + slot);
                    dst[0] = src[0];
                    dst[1] = src[1];
                } else {
                    ++c_overflow;
                }
            }
            slot_cur += __popc(emit_m);
            __syncwarp();
        }
        if (lane == 0) {
            d_tile_tab[2 * chunk] = (uint32_t)chunk_base;
            d_tile_tab[2 * chunk + 1] = (uint32_t)(slot_cur - chunk_base);
        }
    }

    // ---- counters: one warp reduction per counter, one atomic per warp ----------------------------------------------------
    const unsigned r_lines = __reduce_add_sync(0xffffffffu, c_lines), r_kept = __reduce_add_sync(0xffffffffu, c_kept);
    const unsigned r_short = __reduce_add_sync(0xffffffffu, c_short), r_unknown = __reduce_add_sync(0xffffffffu, c_unknown);
    const unsigned r_nnn = __reduce_add_sync(0xffffffffu, c_nnn), r_badpos = __reduce_add_sync(0xffffffffu, c_badpos);
    const unsigned r_slow = __reduce_add_sync(0xffffffffu, c_slow), r_over = __reduce_add_sync(0xffffffffu, c_overflow);
    if (lane == 0) {
        if (r_lines) atomicAdd(&d_counters[MC_C_LINES], (unsigned long long)r_lines);
        if (r_kept) atomicAdd(&d_counters[MC_C_KEPT], (unsigned long long)r_kept);
        if (r_short) atomicAdd(&d_counters[MC_C_SHORT], (unsigned long long)r_short);
        if (r_unknown) atomicAdd(&d_counters[MC_C_UNKNOWN_CONTIG], (unsigned long long)r_unknown);
        if (r_nnn) atomicAdd(&d_counters[MC_C_NNN], (unsigned long long)r_nnn);
        if (r_badpos) atomicAdd(&d_counters[MC_C_BADPOS], (unsigned long long)r_badpos);
        if (r_slow) atomicAdd(&d_counters[MC_C_LONGLINE], (unsigned long long)r_slow);
        if (r_over) atomicAdd(&d_counters[MC_C_OVERFLOW], (unsigned long long)r_over);
    }
}

}  // namespace

extern "C" int64_t mc_num_tiles(int64_t nbytes) { return nbytes <= 0 ? 0 : (nbytes + CHUNK - 1) / CHUNK; }

extern "C" int mc_scan(const uint8_t *d_text, int64_t nbytes, const mc_refindex *ref, int dense, mc_record *d_rec,
                       int64_t rec_cap, uint32_t *d_tile_tab, uint64_t *d_counters, void *stream) {
    MC_REQUIRE(d_text && ref && d_rec && d_tile_tab && d_counters, "null pointer");
    MC_REQUIRE(nbytes >= 0 && rec_cap >= 0, "negative size");
    MC_REQUIRE((reinterpret_cast<uintptr_t>(d_text) & 15) == 0, "d_text must be 16-byte aligned");
    MC_REQUIRE(ref->k >= 1 && ref->k <= MC_MAXK, "k out of range");
    MC_REQUIRE(ref->n_contigs >= 1 && ref->n_contigs < 65535, "contig count out of range");
    MC_REQUIRE(nbytes < (1ll << 47), "chunk too large");
    MC_REQUIRE(rec_cap < (1ll << 32), "record capacity must fit 32 bits");
    const int64_t n_chunks = mc_num_tiles(nbytes);
    if (n_chunks == 0) return MC_OK;
    int dev = 0, sms = 0;
    MC_CUDA_CHECK(cudaGetDevice(&dev));
    MC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = sizeof(WarpSmem) * WARPS;
    MC_CUDA_CHECK(cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_chunks + WARPS - 1) / WARPS;
    const int64_t resident = (int64_t)sms * MC_SCAN_MIN_CTAS;     // persistent: MC_SCAN_MIN_CTAS CTAs of 8 warps per SM
    if (blocks > resident) blocks = resident;
    // 16-byte loads are allowed up to the end of the caller's '\n' padding
    const int64_t text_limit16 = ((nbytes + MC_TEXT_PAD) / 16) * 16;
    k_scan<<<(unsigned)blocks, THREADS, smem, (cudaStream_t)stream>>>(d_text, nbytes, text_limit16, n_chunks, *ref, dense, d_rec,
                                                                     (unsigned long long)rec_cap, d_tile_tab,
                                                                     reinterpret_cast<unsigned long long *>(d_counters));
    MC_LAUNCH_CHECK();
    return MC_OK;
}
