// Stage 1 (K1): eventalign TSV tokeniser + line filter, one CTA per 16 KB text tile.
//
// Replaces the reader / tokeniser / per-line filters of the reference's extract_features
// (extract_contexts.py:140-176): readlines + line.split()[:12], the '<12 fields' drop (:149-152),
// the contig lookup (:154-160), the NNNNNN drop (:167) and the "does this k-mer touch an 'M'"
// test that gates everything after (:176, :242, :269).
//
// Data flow inside a CTA (HBM -> smem once, everything else on-chip):
//   1. the tile (+32 B look-behind, +2 KB look-ahead) is copied to shared memory with 16 B loads;
//   2. SWAR byte classification builds two bit maps per 32 B: non-whitespace (byte > 0x20) and
//      newline (byte == 0x0a); field starts = nonws & ~(nonws << 1);
//   3. line starts inside the tile are compacted into a list (block scan of popcounts);
//   4. one thread per line walks the field-start bits (no byte loop over the ~58 B read name),
//      resolves the contig, parses column 2 and tests the per-position candidate bitmap;
//   5. lines that matter (candidate, successor of a candidate, first kept line of the tile, or all
//      kept lines in dense mode) are compacted and fully parsed (event index, currents as exact
//      decimals -> float64 diff rounded like np.round(x, 4), k-mer equality, read-name span) into
//      32-byte records, written with two 16 B stores.
// Algorithmic HBM traffic: the text itself (once) + ~2 B of records per line in sparse mode.
#include "common.cuh"

namespace {

constexpr int TILE = MC_TILE_BYTES;
constexpr int PRE = 32;
constexpr int OVER = 2048;
constexpr int SMB = PRE + TILE + OVER;      // bytes staged per tile
constexpr int NW = SMB / 32;                // mask words
constexpr int THREADS = 256;
constexpr int WPT = (NW + THREADS - 1) / THREADS;   // mask words per thread (contiguous)
constexpr int LCAP = 1024;                  // line-list capacity per pass
constexpr int ECAP = 1024;                  // emit-list capacity (a tile holds < 16384/23 keepable lines)
static_assert(SMB % 32 == 0, "tile staging must be a whole number of mask words");
static_assert(MC_TEXT_PAD >= OVER + 64, "text padding must cover the look-ahead");

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

struct Smem {
    alignas(16) uint8_t text[SMB + 32];
    uint32_t nonws[NW + 2];
    uint32_t fs[NW + 2];
    uint32_t nl[NW + 2];
    uint16_t lstart[LCAP + 1];
    uint8_t lflag[LCAP];
    uint16_t emit[ECAP];
    int warp_scan[THREADS / 32 + 1];
    unsigned long long cnt[MC_C_COUNT];
    int hint_cid;
    int n_emit;
    int prev_kept_state;   // carry between passes: -1 no kept line yet in tile, 0 last kept not candidate, 1 candidate
    unsigned long long rec_base;
};

__device__ __forceinline__ uint32_t gt20_msb(uint32_t w) { return (((w & 0x7f7f7f7fu) + 0x5f5f5f5fu) | w) & 0x80808080u; }
__device__ __forceinline__ uint32_t eq0a_msb(uint32_t w) {
    uint32_t x = w ^ 0x0a0a0a0au;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// two msb-form words (flags at bits 7,15,23,31) -> 8 flag bits in byte order
__device__ __forceinline__ uint32_t pack8(uint32_t m0, uint32_t m1) { return ((((m0 >> 7) | (m1 >> 3)) * 0x00204081u) >> 21) & 0xFFu; }

__device__ __forceinline__ void masks32(const uint8_t *p, uint32_t &nonws, uint32_t &nl) {
    const uint4 a = *reinterpret_cast<const uint4 *>(p);
    const uint4 b = *reinterpret_cast<const uint4 *>(p + 16);
    nonws = pack8(gt20_msb(a.x), gt20_msb(a.y)) | (pack8(gt20_msb(a.z), gt20_msb(a.w)) << 8) |
            (pack8(gt20_msb(b.x), gt20_msb(b.y)) << 16) | (pack8(gt20_msb(b.z), gt20_msb(b.w)) << 24);
    nl = pack8(eq0a_msb(a.x), eq0a_msb(a.y)) | (pack8(eq0a_msb(a.z), eq0a_msb(a.w)) << 8) |
         (pack8(eq0a_msb(b.x), eq0a_msb(b.y)) << 16) | (pack8(eq0a_msb(b.z), eq0a_msb(b.w)) << 24);
}

// first set bit of mask array m at or after bit position q (returns NW*32 when none)
__device__ __forceinline__ int next_bit(const uint32_t *m, int q) {
    int w = q >> 5;
    uint32_t v = m[w] & (0xFFFFFFFFu << (q & 31));
    while (v == 0u) {
        if (++w >= NW) return NW * 32;
        v = m[w];
    }
    return (w << 5) + __ffs(v) - 1;
}
// first whitespace byte at or after q
__device__ __forceinline__ int token_end(const uint32_t *nonws, int q) {
    int w = q >> 5;
    uint32_t v = ~nonws[w] & (0xFFFFFFFFu << (q & 31));
    while (v == 0u) {
        if (++w >= NW) return NW * 32;
        v = ~nonws[w];
    }
    return (w << 5) + __ffs(v) - 1;
}

struct LineInfo {
    int f[12];       // field start offsets (smem)
    int cid;
    int pos;
    bool kept, cand;
};

__device__ __forceinline__ bool contig_match(const uint8_t *t, int q, const mc_refindex &R, int cid) {
    const int o0 = __ldg(R.d_name_off + cid), o1 = __ldg(R.d_name_off + cid + 1);
    const int L = o1 - o0;
    for (int j = 0; j < L; ++j)
        if (t[q + j] != __ldg(R.d_names + o0 + j)) return false;
    return t[q + L] <= 0x20;
}
// <0, 0, >0 like strcmp(token, name[cid]) with the token ending at the first whitespace
__device__ __forceinline__ int contig_cmp(const uint8_t *t, int q, const mc_refindex &R, int cid) {
    const int o0 = __ldg(R.d_name_off + cid), o1 = __ldg(R.d_name_off + cid + 1);
    const int L = o1 - o0;
    for (int j = 0; j < L; ++j) {
        int a = t[q + j], b = __ldg(R.d_names + o0 + j);
        if (a <= 0x20) return -1;          // token shorter
        if (a != b) return a - b;
    }
    return t[q + L] <= 0x20 ? 0 : 1;
}
// contigs are sorted by name on the host; hint first, then binary search
__device__ __forceinline__ int contig_lookup(const uint8_t *t, int q, const mc_refindex &R, int hint) {
    if (hint >= 0 && contig_match(t, q, R, hint)) return hint;
    int lo = 0, hi = R.n_contigs - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int c = contig_cmp(t, q, R, mid);
        if (c == 0) return mid;
        if (c < 0) hi = mid - 1; else lo = mid + 1;
    }
    return -1;
}

__device__ __forceinline__ bool parse_uint(const uint8_t *t, int q, int &out) {
    int v = 0, nd = 0;
    int c;
    while ((c = t[q]) >= '0' && c <= '9') { v = v * 10 + (c - '0'); ++nd; ++q; if (nd > 9) return false; }
    if (nd == 0 || c > 0x20) return false;
    out = v;
    return true;
}
__device__ __forceinline__ bool parse_int(const uint8_t *t, int q, int &out) {
    bool neg = false;
    if (t[q] == '-') { neg = true; ++q; } else if (t[q] == '+') ++q;
    int v;
    if (!parse_uint(t, q, v)) return false;
    out = neg ? -v : v;
    return true;
}
// plain decimal -> correctly rounded double (mantissa < 2^53, <= 18 digits: one exact division)
__device__ __forceinline__ bool parse_decimal(const uint8_t *t, int q, double &out) {
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
    double v = __ddiv_rn((double)m, c_pow10[nfrac]);
    out = neg ? -v : v;
    return true;
}

// structural parse of one line: field starts, contig, position, candidate bit, kept flag
__device__ __forceinline__ void parse_line(const Smem &S, const mc_refindex &R, int s, int e, int hint, LineInfo &L,
                                           unsigned long long *cnt) {
    L.kept = false; L.cand = false; L.cid = -1; L.pos = 0;
    int w = s >> 5;
    uint32_t m = S.fs[w] & (0xFFFFFFFFu << (s & 31));
#pragma unroll
    for (int f = 0; f < 12; ++f) {
        while (m == 0u && w + 1 < NW) m = S.fs[++w];
        int q = NW * 32;
        if (m != 0u) { q = (w << 5) + __ffs(m) - 1; m &= m - 1u; }
        L.f[f] = q;
    }
    if (L.f[11] >= e) {
        // fewer than 12 fields before the newline (or the line outruns the look-ahead)
        if (cnt) atomicAdd(&cnt[(e >= NW * 32 && L.f[0] < e) ? MC_C_LONGLINE : MC_C_SHORT], 1ull);
        return;
    }
    L.cid = contig_lookup(S.text, L.f[0], R, hint);
    if (L.cid < 0) { if (cnt) atomicAdd(&cnt[MC_C_UNKNOWN_CONTIG], 1ull); return; }
    // model_kmer == 'NNNNNN' (extract_contexts.py:167)
    const uint8_t *t = S.text;
    const int q9 = L.f[9];
    const bool nnn = t[q9] == 'N' && t[q9 + 1] == 'N' && t[q9 + 2] == 'N' && t[q9 + 3] == 'N' && t[q9 + 4] == 'N' &&
                     t[q9 + 5] == 'N' && t[q9 + 6] <= 0x20;
    if (nnn) { if (cnt) atomicAdd(&cnt[MC_C_NNN], 1ull); return; }
    if (!parse_uint(t, L.f[1], L.pos)) { if (cnt) atomicAdd(&cnt[MC_C_BADPOS], 1ull); return; }
    L.kept = true;
    if (L.pos < __ldg(R.d_len + L.cid)) {
        const int64_t g = __ldg(R.d_base + L.cid) + L.pos;
        L.cand = (__ldg(R.d_cand + (g >> 5)) >> (g & 31)) & 1u;
    }
}

__global__ void __launch_bounds__(THREADS, 4)
k_scan(const uint8_t *__restrict__ d_text, int64_t nbytes, int64_t text_limit16, mc_refindex R, int dense,
       mc_record *__restrict__ d_rec, unsigned long long rec_cap, uint32_t *__restrict__ d_tile_tab,
       unsigned long long *__restrict__ d_counters) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t tile = blockIdx.x;
    const int64_t G0 = tile * (int64_t)TILE - PRE;     // global offset of smem byte 0

    if (tid < MC_C_COUNT) S.cnt[tid] = 0ull;
    if (tid == 0) { S.hint_cid = 0; S.n_emit = 0; S.prev_kept_state = -1; }

    // ---- 1. stage the tile --------------------------------------------------------------------------
    for (int i = tid; i < SMB / 16; i += THREADS) {
        const int64_t g = G0 + (int64_t)i * 16;
        uint4 v;
        if (g >= 0 && g < text_limit16) v = __ldcs(reinterpret_cast<const uint4 *>(d_text + g));
        else v = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
        *reinterpret_cast<uint4 *>(S.text + i * 16) = v;
    }
    if (tid < 2) *reinterpret_cast<uint4 *>(S.text + SMB + tid * 16) = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
    __syncthreads();

    // ---- 2. byte classification -> bit maps; 3. line-start list ------------------------------------------
    // thread t owns mask words [t*WPT, t*WPT+WPT)
    uint32_t my_ls[WPT];
    int my_nlines = 0;
    {
        uint32_t prev_nonws_top, prev_nl_top;
        const int w0 = tid * WPT;
        // top bits of the word before my first one (recomputed from the text to avoid a block sync)
        if (w0 > 0 && w0 < NW) {
            uint8_t c = S.text[w0 * 32 - 1];
            prev_nonws_top = c > 0x20;
            prev_nl_top = c == 0x0a;
        } else {
            prev_nonws_top = 0u;
            prev_nl_top = 0u;   // byte before the staged region is never a line start we own (PRE >= 1)
        }
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
            const int w = w0 + j;
            my_ls[j] = 0u;
            if (w < NW) {
                uint32_t nonws, nl;
                masks32(S.text + w * 32, nonws, nl);
                S.nonws[w] = nonws;
                S.nl[w] = nl;
                S.fs[w] = nonws & ~((nonws << 1) | prev_nonws_top);
                uint32_t ls = (nl << 1) | prev_nl_top;          // byte p starts a line iff byte p-1 is '\n'
                // restrict to owned range: PRE <= p < PRE+TILE and global p < nbytes
                const int p0 = w * 32;
                if (p0 + 32 <= PRE || p0 >= PRE + TILE) ls = 0u;
                else {
                    if (p0 < PRE) ls &= 0xFFFFFFFFu << (PRE - p0);
                    if (p0 + 32 > PRE + TILE) ls &= 0xFFFFFFFFu >> (p0 + 32 - (PRE + TILE));
                    const int64_t room = nbytes - (G0 + p0);    // bytes of real text from p0 on
                    if (room <= 0) ls = 0u;
                    else if (room < 32) ls &= (1u << room) - 1u;
                }
                my_ls[j] = ls;
                my_nlines += __popc(ls);
                prev_nonws_top = nonws >> 31;
                prev_nl_top = nl >> 31;
            }
        }
    }
    if (tile == 0 && tid == 0) {
        // the chunk starts at a line start by contract: the look-behind of tile 0 is all '\n'
    }
    int total_lines;
    const int my_first = mc_block_exscan<THREADS>(my_nlines, S.warp_scan, total_lines);
    if (tid == 0) atomicAdd(&S.cnt[MC_C_LINES], (unsigned long long)total_lines);

    // ---- passes over the line list (normally one) ------------------------------------------------------
    for (int pass0 = 0; pass0 < total_lines; pass0 += LCAP) {
        const int n_pass = min(LCAP, total_lines - pass0);
        {
            int idx = my_first - pass0;
#pragma unroll
            for (int j = 0; j < WPT; ++j) {
                uint32_t ls = my_ls[j];
                const int p0 = (tid * WPT + j) * 32;
                while (ls) {
                    const int b = __ffs(ls) - 1;
                    ls &= ls - 1u;
                    if (idx >= 0 && idx < LCAP) S.lstart[idx] = (uint16_t)(p0 + b);
                    ++idx;
                }
            }
        }
        __syncthreads();
        // resolve the contig hint once per pass from the first listed line
        if (tid == 0 && n_pass > 0) {
            const int s = S.lstart[0];
            const int q = next_bit(S.fs, s);
            if (q < NW * 32) {
                int c = contig_lookup(S.text, q, R, S.hint_cid);
                if (c >= 0) S.hint_cid = c;
            }
        }
        __syncthreads();
        const int hint = S.hint_cid;

        // ---- 4. structural parse, one thread per line ----------------------------------------------------
        for (int i = tid; i < n_pass; i += THREADS) {
            const int s = S.lstart[i];
            const int e = (i + 1 < n_pass) ? (int)S.lstart[i + 1] - 1 : next_bit(S.nl, s);
            LineInfo L;
            parse_line(S, R, s, e, hint, L, S.cnt);
            S.lflag[i] = (uint8_t)((L.kept ? 1 : 0) | (L.cand ? 2 : 0));
        }
        __syncthreads();

        // ---- 5a. emit decision + compaction (threads own contiguous lines so order is kept) ----------------
        {
            constexpr int LPT = LCAP / THREADS;
            int my_emit = 0;
            uint32_t emit_bits = 0u;
            int kept_cnt = 0;
            const int carry = S.prev_kept_state;
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                const int i = tid * LPT + j;
                if (i < n_pass) {
                    const int fl = S.lflag[i];
                    if (fl & 1) {
                        ++kept_cnt;
                        bool em = dense || (fl & 2);
                        if (!em) {
                            // previous kept line in this pass, else the carry from earlier passes
                            int st = carry;
                            for (int b = i - 1; b >= 0; --b) {
                                const int fb = S.lflag[b];
                                if (fb & 1) { st = (fb & 2) ? 1 : 0; break; }
                            }
                            em = (st != 0);     // candidate predecessor, or first kept line of the tile (st == -1)
                        }
                        if (em) { emit_bits |= 1u << j; ++my_emit; }
                    }
                }
            }
            int tot_emit;
            int my_off = mc_block_exscan<THREADS>(my_emit, S.warp_scan, tot_emit);
            const int base = S.n_emit;
#pragma unroll
            for (int j = 0; j < LPT; ++j) {
                if (emit_bits & (1u << j)) {
                    const int slot = base + my_off++;
                    if (slot < ECAP) S.emit[slot] = S.lstart[tid * LPT + j];
                }
            }
            if (kept_cnt) atomicAdd(&S.cnt[MC_C_KEPT], (unsigned long long)kept_cnt);
            __syncthreads();
            if (tid == 0) {
                S.n_emit = min(ECAP, base + tot_emit);
                // carry: state of the last kept line of this pass
                for (int b = n_pass - 1; b >= 0; --b) {
                    const int fb = S.lflag[b];
                    if (fb & 1) { S.prev_kept_state = (fb & 2) ? 1 : 0; break; }
                }
            }
            __syncthreads();
        }
    }

    // ---- 5b. allocate record slots for the tile, full parse of the emitted lines -----------------------------
    const int n_emit = S.n_emit;
    if (tid == 0) {
        unsigned long long base = 0ull;
        if (n_emit > 0) base = atomicAdd(&d_counters[MC_C_RECORDS], (unsigned long long)n_emit);
        S.rec_base = base;
        d_tile_tab[2 * tile] = (uint32_t)base;
        d_tile_tab[2 * tile + 1] = (uint32_t)n_emit;
    }
    __syncthreads();
    const unsigned long long rec_base = S.rec_base;
    const int hint = S.hint_cid;
    for (int j = tid; j < n_emit; j += THREADS) {
        const int s = S.emit[j];
        LineInfo L;
        parse_line(S, R, s, NW * 32, hint, L, nullptr);   // already validated: 12 fields exist before the newline
        const uint8_t *t = S.text;
        alignas(16) mc_record r;
        const int64_t goff = G0 + s;
        r.line_lo = (uint32_t)(goff & 0xFFFFFFFFll);
        r.line_hi = (uint16_t)(goff >> 32);
        r.name_off = (uint16_t)(L.f[3] - s);
        r.name_len = (uint16_t)(token_end(S.nonws, L.f[3]) - L.f[3]);
        r.pos = L.pos;
        r.contig = (uint16_t)L.cid;
        uint32_t fl = L.cand ? MC_RF_CAND : 0u;
        int ev_idx = 0;
        if (!parse_int(t, L.f[5], ev_idx)) fl |= MC_RF_BADIDX;
        r.event_idx = ev_idx;
        double ev = 0.0, md = 0.0;
        if (!parse_decimal(t, L.f[6], ev) || !parse_decimal(t, L.f[10], md)) { fl |= MC_RF_BADNUM; r.diff = 0.0; }
        else r.diff = __ddiv_rn(rint(__dmul_rn(__dsub_rn(ev, md), 1e4)), 1e4);      // np.round(ev - model, 4)
        // reference_kmer (col 3) == model_kmer (col 10)
        {
            int a = L.f[2], b = L.f[9];
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
        if (rec_base + j < rec_cap) {
            const uint4 *src = reinterpret_cast<const uint4 *>(&r);
            uint4 *dst = reinterpret_cast<uint4 *>(d_rec + rec_base + j);
            dst[0] = src[0];
            dst[1] = src[1];
        } else {
            atomicAdd(&S.cnt[MC_C_OVERFLOW], 1ull);
        }
    }
    __syncthreads();
    if (tid < MC_C_COUNT && tid != MC_C_RECORDS) {
        const unsigned long long v = S.cnt[tid];
        if (v) atomicAdd(&d_counters[tid], v);
    }
    (void)lane;
}

}  // namespace

extern "C" int64_t mc_num_tiles(int64_t nbytes) { return nbytes <= 0 ? 0 : (nbytes + TILE - 1) / TILE; }

extern "C" int mc_scan(const uint8_t *d_text, int64_t nbytes, const mc_refindex *ref, int dense, mc_record *d_rec,
                       int64_t rec_cap, uint32_t *d_tile_tab, uint64_t *d_counters, void *stream) {
    MC_REQUIRE(d_text && ref && d_rec && d_tile_tab && d_counters, "null pointer");
    MC_REQUIRE(nbytes >= 0 && rec_cap >= 0, "negative size");
    MC_REQUIRE((reinterpret_cast<uintptr_t>(d_text) & 15) == 0, "d_text must be 16-byte aligned");
    MC_REQUIRE(ref->k >= 1 && ref->k <= MC_MAXK, "k out of range");
    MC_REQUIRE(ref->n_contigs >= 1 && ref->n_contigs < 65535, "contig count out of range");
    MC_REQUIRE(nbytes < (1ll << 47), "chunk too large");
    const int64_t n_tiles = mc_num_tiles(nbytes);
    if (n_tiles == 0) return MC_OK;
    MC_REQUIRE(n_tiles < 2147483647ll, "too many tiles");
    static bool attr_set = false;
    if (!attr_set) {
        MC_CUDA_CHECK(cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        attr_set = true;
    }
    // 16-byte loads are allowed up to the end of the caller's '\n' padding
    const int64_t text_limit16 = ((nbytes + MC_TEXT_PAD) / 16) * 16;
    k_scan<<<(unsigned)n_tiles, THREADS, sizeof(Smem), (cudaStream_t)stream>>>(
        d_text, nbytes, text_limit16, *ref, dense, d_rec, (unsigned long long)rec_cap, d_tile_tab,
        reinterpret_cast<unsigned long long *>(d_counters));
    MC_LAUNCH_CHECK();
    return MC_OK;
}
