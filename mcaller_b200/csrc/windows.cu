// Stage 5 (K2): context-window builder -- the state machine of the reference's extract_features
// (extract_contexts.py:169-291) run per read segment over the stage-1 records.
//
// Why this is enough: a kept line whose k-mer holds no 'M' leaves the reference's state machine closed and
// empty (:242-245 / :289-291), and stage 1 records every line that can hold an 'M' on either strand plus the
// kept line that follows it (the line that closes an open window, :179).  Unrecorded lines are therefore no-ops
// and the state machine can run on the records alone.  State never crosses a read boundary except for the one
// still-open window of the previous read, which the first record of the next read closes (:179, `read_name !=
// last_read`); that hand-off is resolved here by looking at the first record of the next segment.
//
// Parallel decomposition: a read is cut into UNITS.  A record that is not a candidate (its k-mer touches no target on
// either strand) leaves the state machine closed and empty (:242-245, :289-291), so the records from the first candidate
// after such a record up to and including the next non-candidate record (the closer) can be processed on their own; the
// only read-level state they need is `last_read` / `first_read_ind` (:161-174), i.e. the read's first line with an 'M',
// which k_first_m finds per read beforehand.  Reads whose first record lies within k of the contig start stay one unit:
// a window at position 0 never closes (`if mpos and`, :179) and keeps its columns across non-candidate lines.
// A block of 256 threads owns 3072 consecutive records, compacts the unit starts among them into shared memory and
// runs one thread per unit; two passes (count rows / write rows) so rows land in file order.  Column sums are
// accumulated in numpy's pairwise order (8 running lanes + sequential tail) so np.mean is reproduced bit for bit.
#include "common.cuh"

namespace {

// Column values live in local memory; the column counts (8 x 8 bits, saturating at 255 -- more than 128 events in a column is
// reported as MC_CE_COLUMN anyway) and the column map (8 x 4 bits) are packed in registers, so the count pass and the
// control flow of the write pass touch no memory.
#ifndef MC_WIN_SMEM_ROWS
#define MC_WIN_SMEM_ROWS 2
#endif
constexpr int SROWS = MC_WIN_SMEM_ROWS;           // first values of every column kept in shared memory (0: all in local memory)
struct ColState {
    double lane[MC_MAXK][8];
    double pend[MC_MAXK][8];
    double *sm;                                   // this thread's [SROWS][MC_MAXK] slice of shared memory, thread-interleaved
    __device__ __forceinline__ double get(int c, int j) const {
        if (SROWS > 0 && j < SROWS) return sm[(j * MC_MAXK + c) * SM_STRIDE];
        return pend[c][j];
    }
    __device__ __forceinline__ void put(int c, int j, double v) {
        if (SROWS > 0 && j < SROWS) sm[(j * MC_MAXK + c) * SM_STRIDE] = v;
        else pend[c][j] = v;
    }
    static constexpr int SM_STRIDE = 256;         // = WIN_THREADS (asserted below): element e of thread t sits at [e][t]
};
__device__ __forceinline__ int cnt_get(unsigned long long cnts, int c) { return (int)((cnts >> (8 * c)) & 0xFFull); }
__device__ __forceinline__ void cnt_inc(unsigned long long &cnts, int c) {
    if (cnt_get(cnts, c) < 255) cnts += 1ull << (8 * c);
}
__device__ __forceinline__ void cnt_clear(unsigned long long &cnts, int c) { cnts &= ~(0xFFull << (8 * c)); }
__device__ __forceinline__ int map_get(uint32_t mp, int c) { return (int)((mp >> (4 * c)) & 0xFu); }

__device__ __forceinline__ void col_push(ColState &C, unsigned long long &cnts, int c, double v) {
    const int n = cnt_get(cnts, c);
    const int j = n & 7;
    C.put(c, j, v);
    cnt_inc(cnts, c);
    if (j == 7) {
        if (n == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) C.lane[c][i] = C.get(c, i);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) C.lane[c][i] = __dadd_rn(C.lane[c][i], C.get(c, i));
        }
    }
}

// numpy add.reduce pairwise order for n <= 128 (see oracle/mcaller_oracle.c np_pairwise), then / n
__device__ __forceinline__ double col_mean(const ColState &C, unsigned long long cnts, int c) {
    const int n = cnt_get(cnts, c);
    double res;
    if (n < 8) {
        res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, C.get(c, i));
    } else {
        const double *r = C.lane[c];
        res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                        __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (int i = 0; i < (n & 7); ++i) res = __dadd_rn(res, C.get(c, i));
    }
    return __ddiv_rn(res, (double)n);
}

__device__ __forceinline__ int64_t rec_line(const mc_record &r) { return ((int64_t)r.line_hi << 32) | (int64_t)r.line_lo; }

__device__ __forceinline__ mc_record load_rec(const mc_record *p) {
    mc_record r;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(s), b = __ldg(s + 1);
    uint4 *d = reinterpret_cast<uint4 *>(&r);
    d[0] = a;
    d[1] = b;
    return r;
}

__device__ __forceinline__ uint8_t comp_base(uint8_t c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
        case 'N': return 'N';
        default: return 0;
    }
}

#ifndef MC_WIN_ITEMS
#define MC_WIN_ITEMS 12
#endif
constexpr int WIN_THREADS = 256, WIN_ITEMS = MC_WIN_ITEMS, WIN_RECS = WIN_THREADS * WIN_ITEMS;   // 3072 records per block: ~230 units, one per thread
static_assert(WIN_RECS < 65536, "unit / read counts of a block are packed in 16 bits");
static_assert(ColState::SM_STRIDE == WIN_THREADS, "shared-memory column slices are interleaved by thread");

// first line of each read with an 'M' under the per-line strand guess (:169-176): record index and event index
__global__ void __launch_bounds__(128) k_first_m(const mc_record *__restrict__ rec, const uint32_t *__restrict__ seg_start, int64_t seg_cap,
                                                const unsigned long long *__restrict__ d_nseg, mc_refindex R,
                                                uint32_t *__restrict__ first_idx, int32_t *__restrict__ first_ind) {
    const int64_t n_seg = mc_dev_count(d_nseg, seg_cap);
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    const uint32_t b = seg_start[seg], e = seg_start[seg + 1];
    uint32_t fi = 0xFFFFFFFFu;
    int32_t ev = 0;
    for (uint32_t i = b; i < e; ++i) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(rec + i)), c = __ldg(reinterpret_cast<const uint4 *>(rec + i) + 1);
        const uint32_t fl = c.w & 0xFFu;
        if (!(fl & MC_RF_CAND)) continue;
        const int pos = (int)a.z, cid = (int)(c.z >> 16);
        const int rev = !(fl & MC_RF_EQ);
        uint32_t bits = 0u;
        if (pos < __ldg(R.d_len + cid)) bits = mc_kmer_bits(rev ? R.d_site_rev : R.d_site_fwd, __ldg(R.d_base + cid) + pos, R.k);
        if (bits) { fi = i; ev = (int32_t)a.w; break; }
    }
    first_idx[seg] = fi;
    first_ind[seg] = ev;
}

template <bool WRITE>
__global__ void __launch_bounds__(WIN_THREADS)
k_windows(const mc_record *__restrict__ rec, int64_t rec_cap, const unsigned long long *__restrict__ d_n_records,
          const uint32_t *__restrict__ seg_start, int64_t seg_cap, const unsigned long long *__restrict__ d_nseg,
          const double *__restrict__ seg_qual, const uint32_t *__restrict__ first_idx, const int32_t *__restrict__ first_ind_arr,
          mc_refindex R, int skip_thresh, double qual_thresh, int two_models, mc_call *__restrict__ calls,
          unsigned long long call_cap, uint32_t *__restrict__ unit_cnt /* [n_records], written at unit starts */,
          uint32_t *__restrict__ blk_tot, const uint32_t *__restrict__ blk_off) {
    __shared__ uint32_t s_unit[WIN_RECS];         // record index of each unit start of the block, in record order
    __shared__ uint32_t s_off[WIN_RECS];          // write pass: row offset of each unit
    __shared__ uint16_t s_seg[WIN_RECS];          // read segment of each unit, relative to the block's first record's
    __shared__ int s_warp[WIN_THREADS / 32 + 1];
    __shared__ int64_t s_seg0;
    extern __shared__ __align__(16) double s_cols[];              // write pass: [SROWS][MC_MAXK][WIN_THREADS] first values of each column
    const int k = R.k;
    // the record / segment counts live on the device; the grid was sized for rec_cap records
    const int64_t n_records = mc_dev_count(d_n_records, rec_cap), n_seg = mc_dev_count(d_nseg, seg_cap);
    if ((int64_t)blockIdx.x * WIN_RECS >= n_records || n_seg <= 0) {       // block-uniform
        if (!WRITE && threadIdx.x == 0) blk_tot[blockIdx.x] = 0u;
        return;
    }
    // ---- unit starts among this block's records, compacted in record order ------------------------------------------------
    int nu;
    {
        const int64_t base = (int64_t)blockIdx.x * WIN_RECS;
        if (threadIdx.x == 0) {                   // segment of the block's first record: last segment starting at or before it
            int64_t lo = 0, hi = n_seg - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if ((int64_t)__ldg(seg_start + mid) <= base) lo = mid; else hi = mid - 1;
            }
            s_seg0 = lo;
        }
        const int64_t i0 = base + (int64_t)threadIdx.x * WIN_ITEMS;
        uint32_t prev_cand = 0u;
        if (i0 > 0 && i0 <= n_records) prev_cand = __ldg(reinterpret_cast<const uint4 *>(rec + i0 - 1) + 1).w & MC_RF_CAND;
        uint32_t umask = 0u, nmask = 0u;          // unit starts / read starts among this thread's records
#pragma unroll
        for (int j = 0; j < WIN_ITEMS; ++j) {
            const int64_t i = i0 + j;
            if (i < n_records) {
                const uint32_t fl = __ldg(reinterpret_cast<const uint4 *>(rec + i) + 1).w & 0xFFu;
                const uint32_t cand = fl & MC_RF_CAND;
                bool u;
                if (i == 0 || (fl & MC_RF_NEWREAD)) {
                    if (i != base) nmask |= 1u << j;
                    u = cand != 0u;
                    if (!u) u = (int)__ldg(reinterpret_cast<const uint4 *>(rec + i)).z < k;      // whole-read unit (see above)
                } else {
                    u = cand != 0u && prev_cand == 0u;
                }
                if (u) umask |= 1u << j;
                prev_cand = cand;
            }
        }
        int total;
        const int off = mc_block_exscan<WIN_THREADS>(__popc(umask) | (__popc(nmask) << 16), s_warp, total);
        int ou = off & 0xFFFF;
        const int on = off >> 16;
        for (uint32_t m = umask; m; m &= m - 1u) {
            const int j = __ffs(m) - 1;
            s_unit[ou] = (uint32_t)(i0 + j);
            s_seg[ou] = (uint16_t)(on + __popc(nmask & ((2u << j) - 1u)));
            ++ou;
        }
        nu = total & 0xFFFF;
        __syncthreads();
    }
    if (WRITE) {
        // row offsets of the block's units: counts of the first pass -> exclusive prefix in shared memory
        for (int u = threadIdx.x; u < WIN_RECS; u += WIN_THREADS) s_off[u] = u < nu ? unit_cnt[s_unit[u]] : 0u;
        __syncthreads();
        uint32_t v[WIN_ITEMS];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < WIN_ITEMS; ++j) { v[j] = s_off[threadIdx.x * WIN_ITEMS + j]; sum += (int)v[j]; }
        int total;
        uint32_t run = blk_off[blockIdx.x] + (uint32_t)mc_block_exscan<WIN_THREADS>(sum, s_warp, total);
#pragma unroll
        for (int j = 0; j < WIN_ITEMS; ++j) { s_off[threadIdx.x * WIN_ITEMS + j] = run; run += v[j]; }
        __syncthreads();
    }
    int my_rows = 0;
    for (int un = threadIdx.x; un < nu; un += WIN_THREADS) {
    const uint32_t b = s_unit[un];
    const int64_t seg = s_seg0 + s_seg[un];
    const uint32_t seg_b = __ldg(seg_start + seg), e_read = __ldg(seg_start + seg + 1);
    const bool whole = (int)__ldg(reinterpret_cast<const uint4 *>(rec + seg_b)).z < k;     // the read is one unit
    uint32_t n_out = 0;
    const double myq = seg_qual[seg];
    if (myq < qual_thresh || (whole && b != seg_b)) {          // whole read dropped (:167) / unit covered by the whole-read unit
        if (!WRITE) unit_cnt[b] = 0u;
        continue;
    }
    uint32_t e = e_read;
    uint32_t out_pos = WRITE ? s_off[un] : 0u;
    const uint32_t fidx = __ldg(first_idx + seg);

    alignas(16) ColState C;
    C.sm = WRITE ? s_cols + threadIdx.x : nullptr;
    unsigned long long cnts = 0ull;                // events per column slot
    uint32_t mp = 0x76543210u;                     // window column -> column slot (the multi-M carry permutes it)
    bool started = b > fidx;     // read_name == last_read  (this read already had a line with 'M')
    bool has_mpos = false;
    int mpos = 0, first_ind = started ? __ldg(first_ind_arr + seg) : 0, last_rev = 0, last_cid = 0;
    uint32_t name_rec = b;
    uint32_t sticky_err = 0u;

#define MPOS_TRUTHY (has_mpos && mpos != 0)

    // emits the row(s) for the open window; `close_idx` = ordered index of the closing record (or ~0u)
    auto emit_window = [&](uint32_t close_idx, int chrom_cid) {
        int n_empty = 0;
        for (int c = 0; c < k; ++c) n_empty += (cnt_get(cnts, map_get(mp, c)) == 0);
        if (WRITE) {
            if (out_pos < call_cap) {
                mc_call &o = calls[out_pos];
                const mc_record nr = load_rec(rec + name_rec);
                o.read_off = rec_line(nr) + nr.name_off;
                o.read_len = nr.name_len;
                o.prob = 0.0;
                o.mpos = mpos;
                o.close_rec = close_idx;
                o.win_contig = (uint16_t)last_cid;
                o.chrom_contig = (uint16_t)(chrom_cid < 0 ? 0xFFFF : chrom_cid);
                o.rev = (uint8_t)last_rev;
                o.n_empty = (uint8_t)n_empty;
                o.label = 0;
                o.pad0 = 0;
                o.seg = (uint32_t)seg;
                o.pad1 = 0;
                uint32_t err = sticky_err;
                uint32_t empty_mask = 0u;
                const int64_t g = __ldg(R.d_base + last_cid) + mpos;
                // dense site slot = rank of the target among the strand's sites
                {
                    const uint32_t *bm = last_rev ? R.d_site_rev : R.d_site_fwd;
                    const uint32_t *rk = last_rev ? R.d_rank_rev : R.d_rank_fwd;
                    const uint32_t wbits = __ldg(bm + (g >> 5));
                    o.site = (int32_t)(__ldg(rk + (g >> 5)) + __popc(wbits & ((1u << (g & 31)) - 1u)));
                }
                if (n_empty <= skip_thresh) {
                    o.kind = MC_CALL;
                    for (int c = 0; c < k; ++c) {                 // :186-188 (forward reads are flipped)
                        const int src = map_get(mp, last_rev ? c : (k - 1 - c));
                        if (cnt_get(cnts, src) == 0) { o.feat[c] = 0.0; empty_mask |= 1u << c; }
                        else {
                            if (cnt_get(cnts, src) > 128) err |= MC_CE_COLUMN;
                            o.feat[c] = col_mean(C, cnts, src);
                        }
                    }
                    o.feat[k] = myq;                              // :189-193
                    for (int c = k + 1; c <= MC_MAXK; ++c) o.feat[c] = 0.0;
                    // context bounds (:194-195) and base after the target (:197, base_models :99-106)
                    const int clen = __ldg(R.d_len + last_cid);
                    uint8_t nextb = 0;
                    if (mpos - k + 1 < 0 || mpos + k > clen) err |= MC_CE_CONTEXT;
                    else if (!last_rev) {
                        const int64_t gn = g + 1;
                        nextb = ((__ldg(R.d_site_fwd + (gn >> 5)) >> (gn & 31)) & 1u) ? 'M' : __ldg(R.d_bases + gn);
                    } else {
                        const int64_t gn = g - 1;
                        nextb = ((__ldg(R.d_site_rev + (gn >> 5)) >> (gn & 31)) & 1u) ? 'M' : comp_base(__ldg(R.d_bases + gn));
                    }
                    if (!(err & MC_CE_CONTEXT) &&
                        !(nextb == 'A' || nextb == 'C' || nextb == 'G' || nextb == 'T' || nextb == 'M'))
                        err |= MC_CE_MODELKEY;
                    o.model_sel = (uint8_t)((two_models && nextb == 'G') ? 1 : 0);
                } else {
                    o.kind = MC_TOO_MANY_SKIPS;                   // :238-239
                    for (int c = 0; c <= MC_MAXK; ++c) o.feat[c] = 0.0;
                    o.model_sel = 0;
                }
                o.empty_mask = (uint8_t)empty_mask;
                o.err = (uint8_t)err;
            }
            ++out_pos;
        }
        ++n_out;
    };
    auto emit_multi = [&]() {
        if (WRITE) {
            if (out_pos < call_cap) {
                mc_call &o = calls[out_pos];
                const mc_record nr = load_rec(rec + name_rec);
                o.read_off = rec_line(nr) + nr.name_off;
                o.read_len = nr.name_len;
                o.prob = 0.0;
                for (int c = 0; c <= MC_MAXK; ++c) o.feat[c] = 0.0;
                o.mpos = mpos;
                o.site = -1;
                o.close_rec = 0u;
                o.win_contig = (uint16_t)last_cid;
                o.chrom_contig = (uint16_t)last_cid;
                o.kind = MC_MULTI_M;
                o.rev = (uint8_t)last_rev;
                o.n_empty = 0; o.empty_mask = 0; o.model_sel = 0; o.label = 0; o.err = 0; o.pad0 = 0;
                o.seg = (uint32_t)seg;
                o.pad1 = 0;
            }
            ++out_pos;
        }
        ++n_out;
    };
    auto reset_cols = [&]() { cnts = 0ull; };

    // The row-producing code (column means, divisions, context look-ups) is by far the heaviest path and a lane needs it
    // only once per ~20 records.  Run in rounds so the warp executes it together: each lane advances through its records
    // until it reaches a window close (or its end-of-read hand-off), then all lanes that have one emit, then the next round.
    uint32_t i = b;
    mc_record r = (b < e) ? load_rec(rec + b) : mc_record();
    mc_record r_next = (b + 1 < e) ? load_rec(rec + b + 1) : mc_record();
    auto advance = [&]() {
        ++i;
        r = r_next;
        if (i + 1 < e) r_next = load_rec(rec + i + 1);              // overlap the next record's latency with this one's work
        // a lane streams its own 32-byte records: pull the 128-byte line 16 records ahead into L2 (L1 is left to the
        // column state, which lives in local memory).  Count pass only: in the write pass the prefetch was measured slower.
        if (!WRITE && (i & 3u) == 0u && (int64_t)i + 16 < n_records) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + i + 16));
    };
    int rev = 0, first_m = -1, cid = 0, pos = 0;
    bool pending_close = false, handoff_done = false;
    // part of a record's processing after the close decision: feed or reset (:269-291)
    auto feed = [&]() {
        if (first_m >= 0) {                                        // :269-287
            if (MPOS_TRUTHY && rev != last_rev) has_mpos = false;  // columns kept (:276-277)
            if (!MPOS_TRUTHY) { has_mpos = true; mpos = pos + first_m; }
            started = true;
            last_rev = rev;
            last_cid = cid;
            name_rec = i;
            if (r.flags & MC_RF_BADNUM) sticky_err |= MC_CE_BADNUM;
            if (WRITE) col_push(C, cnts, map_get(mp, first_m), r.diff);
            else cnt_inc(cnts, map_get(mp, first_m));
        } else if (MPOS_TRUTHY) {                                  // :289-291
            has_mpos = false;
            reset_cols();
        }
    };
    for (;;) {
        // ---- phase 1: advance to the next emission point ---------------------------------------------------------------
        while (!pending_close && i < e) {
            if (!whole && !(r.flags & MC_RF_CAND)) e = i + 1;      // the closer ends the unit
            const bool same_read = started;
            if (!same_read) {                                      // :161-162
                first_ind = r.event_idx;
                if (r.flags & MC_RF_BADIDX) sticky_err |= MC_CE_BADNUM;
            }
            if (!same_read) rev = !(r.flags & MC_RF_EQ);           // :169-174
            else {
                if (r.flags & MC_RF_BADIDX) sticky_err |= MC_CE_BADNUM;
                rev = !(r.event_idx > first_ind);
            }
            cid = r.contig;
            pos = r.pos;
            uint32_t bits = 0u;                                    // 'M's of meth_ref[pos:pos+k] (:176)
            if (pos < __ldg(R.d_len + cid)) bits = mc_kmer_bits(rev ? R.d_site_rev : R.d_site_fwd, __ldg(R.d_base + cid) + pos, k);
            first_m = bits ? (__ffs(bits) - 1) : -1;
            if (MPOS_TRUTHY && pos >= mpos + 1 && same_read) {     // :179 (the other-read case is the segment hand-off below)
                pending_close = true;
                break;
            }
            feed();
            advance();
        }
        // hand-off: a window still open at the end of the read is closed by the next kept line of the file, i.e. the
        // first record of the next segment that passes the quality filter (:179, read_name != last_read)
        const bool pending_handoff = !pending_close && i >= e && !handoff_done && MPOS_TRUTHY;
        if (!pending_close && !pending_handoff) break;
        // ---- phase 2: the lanes that reached an emission point emit together --------------------------------------------
        uint32_t close_idx = i;
        int chrom = cid;
        if (pending_handoff) {
            int64_t j = seg + 1;
            while (j < n_seg && seg_qual[j] < qual_thresh) ++j;
            if (j < n_seg) {
                close_idx = seg_start[j];
                chrom = load_rec(rec + close_idx).contig;
            } else {
                close_idx = 0xFFFFFFFFu;                           // pending: resolved by the next chunk, dropped at EOF
                chrom = -1;
            }
        }
        emit_window(close_idx, chrom);
        if (pending_handoff) {
            handoff_done = true;
            continue;
        }
        if (first_m < 0 || pos > mpos + skip_thresh + 1) {         // :242-245
            reset_cols();
            has_mpos = false;
        } else {                                                   // :246-256 multi-M carry
            if (first_m != 0) emit_multi();
            const int last_mpos = mpos;
            mpos = pos + first_m;
            int sp = mpos - last_mpos;
            if (sp > k) sp = k;
            if (sp <= 0) { sticky_err |= MC_CE_SPACING; sp = k; }
            uint32_t nmp = 0u;
            for (int c = 0; c < k; ++c) {
                int slot;
                if (c < sp) { slot = map_get(mp, k - sp + c); cnt_clear(cnts, slot); }
                else slot = map_get(mp, c - sp);
                nmp |= (uint32_t)slot << (4 * c);
            }
            mp = nmp | (mp & ~((k < 8) ? ((1u << (4 * k)) - 1u) : 0xFFFFFFFFu));
        }
        pending_close = false;
        feed();
        advance();
    }
    if (!WRITE) unit_cnt[b] = n_out;
    my_rows += (int)n_out;
#undef MPOS_TRUTHY
    }
    if (!WRITE) {
        __syncthreads();
        int total;
        (void)mc_block_exscan<WIN_THREADS>(my_rows, s_warp, total);
        if (threadIdx.x == 0) blk_tot[blockIdx.x] = (uint32_t)total;
    }
}

__global__ void k_check_cap(const unsigned long long *d_ncalls, unsigned long long cap, unsigned long long *d_flag) {
    if (d_ncalls[0] > cap) *d_flag = 1ull;
}

}  // namespace

extern "C" int mc_build_windows(const mc_record *d_rec, const uint64_t *d_n_records, int64_t rec_cap, const uint32_t *d_seg_start,
                                const uint64_t *d_nseg, int64_t seg_cap, const double *d_seg_qual, const mc_refindex *ref,
                                int skip_thresh, double qual_thresh, int two_models, mc_call *d_calls, int64_t call_cap,
                                uint32_t *d_seg_count, uint64_t *d_ncalls, void *d_ws, void *stream) {
    MC_REQUIRE(d_rec && d_n_records && d_seg_start && d_nseg && d_seg_qual && ref && d_calls && d_seg_count && d_ncalls && d_ws,
               "null pointer");
    MC_REQUIRE(ref->k >= 1 && ref->k <= MC_MAXK, "k out of range");
    MC_REQUIRE(skip_thresh >= 0, "skip_thresh must be >= 0");
    cudaStream_t st = (cudaStream_t)stream;
    MC_CUDA_CHECK(cudaMemsetAsync(d_ncalls, 0, 16, st));
    if (seg_cap <= 0 || rec_cap <= 0) return MC_OK;
    MC_REQUIRE(rec_cap < (1ll << 32), "record count must fit 32 bits");
    if (seg_cap > rec_cap) seg_cap = rec_cap;                       // a segment holds at least one record
    const unsigned long long *dn = reinterpret_cast<const unsigned long long *>(d_n_records);
    const unsigned long long *ds = reinterpret_cast<const unsigned long long *>(d_nseg);
    // workspace: [unit_cnt: u32 rec_cap][first_ind: i32 seg_cap][blk_tot, blk_off: u32 nb each][scan sums]
    const int64_t nb = (rec_cap + WIN_RECS - 1) / WIN_RECS;
    auto up = [](int64_t bytes) { return ((bytes + 255) / 256) * 256; };
    uint8_t *w = reinterpret_cast<uint8_t *>(d_ws);
    uint32_t *unit_cnt = reinterpret_cast<uint32_t *>(w);
    int32_t *first_ind = reinterpret_cast<int32_t *>(w + up(rec_cap * 4));
    uint32_t *blk_tot = reinterpret_cast<uint32_t *>(w + up(rec_cap * 4) + up(seg_cap * 4));
    uint32_t *blk_off = blk_tot + nb;
    void *scan_ws = w + up(rec_cap * 4) + up(seg_cap * 4) + up(2 * nb * 4);
    uint32_t *first_idx = d_seg_count;                             // caller's seg_cap-sized scratch
    k_first_m<<<(unsigned)((seg_cap + 127) / 128), 128, 0, st>>>(d_rec, d_seg_start, seg_cap, ds, *ref, first_idx, first_ind);
    MC_LAUNCH_CHECK();
    k_windows<false><<<(unsigned)nb, WIN_THREADS, 0, st>>>(d_rec, rec_cap, dn, d_seg_start, seg_cap, ds, d_seg_qual, first_idx, first_ind,
                                                          *ref, skip_thresh, qual_thresh, two_models, d_calls,
                                                          (unsigned long long)call_cap, unit_cnt, blk_tot, nullptr);
    MC_LAUNCH_CHECK();
    int rc = mc_exscan_u32(blk_tot, blk_off, nb, d_ncalls, scan_ws, st);
    if (rc) return rc;
    constexpr size_t cols_bytes = sizeof(double) * SROWS * MC_MAXK * WIN_THREADS;
    MC_CUDA_CHECK(cudaFuncSetAttribute(k_windows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_bytes));
    k_windows<true><<<(unsigned)nb, WIN_THREADS, cols_bytes, st>>>(d_rec, rec_cap, dn, d_seg_start, seg_cap, ds, d_seg_qual, first_idx,
                                                                  first_ind, *ref, skip_thresh, qual_thresh, two_models, d_calls,
                                                                  (unsigned long long)call_cap, unit_cnt, nullptr, blk_off);
    MC_LAUNCH_CHECK();
    k_check_cap<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long *>(d_ncalls), (unsigned long long)call_cap,
                                 reinterpret_cast<unsigned long long *>(d_ncalls) + 1);
    MC_LAUNCH_CHECK();
    return MC_OK;
}
