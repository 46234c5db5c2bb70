// make_bed drop-in: per-position aggregation straight from a `.diffs.<k>` text file (make_bed.py:75-98).
// Rows are tab-split exactly like the reference (8 fields, or 7 for the old format); the locus key is
// (chrom, pos, context, strand); depth / methylated counts are integer atomics in an open-addressing table and the
// first-seen order of make_bed.py:134 is recovered from the smallest line offset per key.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_diffs_aggregate(const uint8_t *__restrict__ text, int64_t nbytes, mc_locus_entry *__restrict__ table, unsigned long long mask,
                  unsigned long long *__restrict__ counters) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nbytes) return;
    if (p > 0 && __ldg(text + p - 1) != '\n') return;          // not a line start
    // split on '\t' up to the newline (the newline stays inside the last field, like Python's split('\t'))
    int64_t fs[9];
    int nf = 0;
    fs[0] = p;
    int64_t q = p;
    for (; q < nbytes; ++q) {
        const uint8_t c = __ldg(text + q);
        if (c == '\n') break;
        if (c == '\t') { ++nf; if (nf < 9) fs[nf] = q + 1; }
    }
    ++nf;
    const int64_t end = q;                                       // position of '\n' (or nbytes)
    atomicAdd(&counters[0], 1ull);                               // lines
    if (nf != 8 && nf != 7) { atomicAdd(&counters[1], 1ull); return; }   // reference: unpack error
    const int64_t c0 = fs[0], c0e = fs[1] - 1, p0 = fs[2], p0e = fs[3] - 1, x0 = fs[3], x0e = fs[4] - 1, s0 = fs[5], s0e = fs[6] - 1;
    const int64_t l0 = fs[6];
    const int64_t xl = x0e - x0;
    if (xl <= 0 || __ldg(text + x0 + xl / 2) != 'M') { atomicAdd(&counters[2], 1ull); return; }   // :84
    unsigned long long h = 14695981039346656037ull;
    for (int64_t i = c0; i < c0e; ++i) h = (h ^ __ldg(text + i)) * 1099511628211ull;
    h = (h ^ 9ull) * 1099511628211ull;
    for (int64_t i = p0; i < p0e; ++i) h = (h ^ __ldg(text + i)) * 1099511628211ull;
    h = (h ^ 9ull) * 1099511628211ull;
    for (int64_t i = x0; i < x0e; ++i) h = (h ^ __ldg(text + i)) * 1099511628211ull;
    h = (h ^ 9ull) * 1099511628211ull;
    for (int64_t i = s0; i < s0e; ++i) h = (h ^ __ldg(text + i)) * 1099511628211ull;
    if (h == 0ull) h = 1ull;
    const bool is_m = (l0 < end || nf == 8) && __ldg(text + l0) == 'm';      // label[0] == 'm' (:93)
    unsigned long long slot = h & mask;
    for (unsigned long long probe = 0; probe <= mask; ++probe) {
        unsigned long long cur = table[slot].hash;
        if (cur == 0ull) {
            cur = atomicCAS(&table[slot].hash, 0ull, h);
            if (cur == 0ull) cur = h;
        }
        if (cur == h) {
            atomicAdd(&table[slot].depth, 1u);
            if (is_m) atomicAdd(&table[slot].meth, 1u);
            atomicMin(&table[slot].first_off, (unsigned long long)p);
            return;
        }
        slot = (slot + 1) & mask;
    }
    atomicAdd(&counters[3], 1ull);                               // table full
}

}  // namespace

extern "C" int mc_diffs_aggregate(const uint8_t *d_text, int64_t nbytes, mc_locus_entry *d_table, int64_t table_size,
                                  uint64_t *d_counters, void *stream) {
    MC_REQUIRE(d_text && d_table && d_counters, "null pointer");
    MC_REQUIRE(table_size > 0 && (table_size & (table_size - 1)) == 0, "table size must be a power of two");
    if (nbytes <= 0) return MC_OK;
    k_diffs_aggregate<<<(unsigned)((nbytes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_text, nbytes, d_table, (unsigned long long)(table_size - 1), reinterpret_cast<unsigned long long *>(d_counters));
    MC_LAUNCH_CHECK();
    return MC_OK;
}
