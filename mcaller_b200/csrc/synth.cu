// Synthetic `nanopolish eventalign` generator on the device (bench / test tooling, not part of the hot path).
// Bit-identical to mcaller_b200/synth.py: every value is a pure function of (seed, read index, position, event)
// through the same splitmix64-based hash, integer arithmetic only.  It exists so bench.py can fill HBM with the
// BASELINE configuration (100k reads ~ 48 GB of TSV) without generating or copying anything on the host.
// One thread per read, two passes: sizes (-> exclusive scan by the caller) and write.
#include "common.cuh"

namespace {

enum { S_REF = 0, S_MODEL, S_LEN, S_START, S_STRAND, S_NAME, S_QUAL, S_POS, S_EV, S_EV2, S_E0, S_METH, S_NEXTRA };
constexpr unsigned long long GOLD = 0x9E3779B97F4A7C15ull;
constexpr int KM = 6;

__host__ __device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    unsigned long long z = x + GOLD;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ unsigned long long stream_seed(unsigned long long seed, int s) { return mix(seed * 64ull + (unsigned long long)s); }
__host__ __device__ __forceinline__ unsigned long long H(unsigned long long s, unsigned long long a, unsigned long long b) { return mix(mix(s + a) + b); }

__constant__ int c_meth_off[6] = {150, -250, 400, -350, 300, 120};
__constant__ char c_suffix[] = "_Basecall_1D_template";

template <bool WRITE>
struct Sink {
    uint8_t *p;
    unsigned long long n;
    __device__ __forceinline__ void put(uint8_t c) {
        if (WRITE) p[n] = c;
        ++n;
    }
    __device__ __forceinline__ void put_uint(unsigned v) {
        char tmp[12];
        int k = 0;
        do { tmp[k++] = (char)('0' + v % 10u); v /= 10u; } while (v);
        while (k) put((uint8_t)tmp[--k]);
    }
    __device__ __forceinline__ void put_fixed(unsigned v, int digits) {   // zero padded
        char tmp[12];
        for (int k = 0; k < digits; ++k) { tmp[k] = (char)('0' + v % 10u); v /= 10u; }
        for (int k = digits - 1; k >= 0; --k) put((uint8_t)tmp[k]);
    }
    __device__ __forceinline__ void put_centi(int v) {                    // "%.2f" of v/100
        if (v < 0) { put('-'); v = -v; }
        put_uint((unsigned)v / 100u);
        put('.');
        put_fixed((unsigned)v % 100u, 2);
    }
};

struct ReadMeta {
    int ci, start, length, rev, e0;
};

__device__ __forceinline__ ReadMeta read_meta(const mc_synth_spec &S, long long i) {
    ReadMeta m;
    int ci = 0;
    while (!(S.d_read_bounds[ci] <= i && i < S.d_read_bounds[ci + 1])) ++ci;
    const long long clen = S.d_contig_len[ci];
    const long long n_c = S.d_read_bounds[ci + 1] - S.d_read_bounds[ci];
    const long long i_c = i - S.d_read_bounds[ci];
    m.ci = ci;
    m.length = S.len_min + (int)(H(stream_seed(S.seed, S_LEN), (unsigned long long)i, 0) % (unsigned long long)(S.len_max - S.len_min + 1));
    const long long span = clen - 2 * S.margin - S.len_max - KM;
    long long stride = span / n_c;
    if (stride < 1) stride = 1;
    long long st = (i_c * span) / n_c + (long long)(H(stream_seed(S.seed, S_START), (unsigned long long)i, 0) % (unsigned long long)stride);
    if (st > span) st = span;
    m.start = S.margin + (int)st;
    m.rev = (int)(H(stream_seed(S.seed, S_STRAND), (unsigned long long)i, 0) & 1ull);
    m.e0 = (int)(H(stream_seed(S.seed, S_E0), (unsigned long long)i, 0) % 50ull);
    return m;
}

__device__ __forceinline__ int events_at(const mc_synth_spec &S, unsigned long long s_pos, unsigned long long s_extra, int p) {
    const unsigned long long hp = H(s_pos, (unsigned long long)p, 0);
    if ((int)(hp & 0xFFFFull) < S.p_skip) return 0;
    int nev = 1;
    bool alive = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        alive = alive && (((hp >> (16 + 8 * j)) & 0xFFull) < 123ull);
        nev += alive ? 1 : 0;
    }
    if (nev == 7) nev = 7 + (int)(H(s_extra, (unsigned long long)p, 0) % 14ull);
    return nev;
}

__device__ __forceinline__ int base_code(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

template <bool WRITE>
__device__ unsigned long long gen_read(const mc_synth_spec &S, long long i, uint8_t *dst) {
    const ReadMeta m = read_meta(S, i);
    const unsigned long long ui = (unsigned long long)i;
    const unsigned long long s_pos = stream_seed(S.seed, S_POS) + ui * GOLD;
    const unsigned long long s_extra = stream_seed(S.seed, S_NEXTRA) + ui * GOLD;
    const unsigned long long s_ev = stream_seed(S.seed, S_EV) + ui * GOLD;
    const unsigned long long s_ev2 = stream_seed(S.seed, S_EV2) + ui * GOLD;
    // read name: 32 hex digits as 8-4-4-4-12 + suffix
    char name[64];
    int nlen = 0;
    {
        const unsigned long long a = H(stream_seed(S.seed, S_NAME), ui, 0), b = H(stream_seed(S.seed, S_NAME), ui, 1);
        for (int d = 0; d < 32; ++d) {
            const unsigned long long src = d < 16 ? a : b;
            const int nib = (int)((src >> (60 - 4 * (d & 15))) & 0xFull);
            if (d == 8 || d == 12 || d == 16 || d == 20) name[nlen++] = '-';
            name[nlen++] = (char)(nib < 10 ? '0' + nib : 'a' + nib - 10);
        }
        for (int j = 0; c_suffix[j]; ++j) name[nlen++] = c_suffix[j];
    }
    int total = 0;
    if (m.rev) {
        for (int pi = 0; pi < m.length; ++pi) total += events_at(S, s_pos, s_extra, m.start + pi);
    }
    const uint8_t *cname = S.d_names + S.d_name_off[m.ci];
    const int cname_len = S.d_name_off[m.ci + 1] - S.d_name_off[m.ci];
    const uint8_t *g = S.d_genome + S.d_gbase[m.ci];
    const uint32_t *mbm = S.meth ? (m.rev ? S.d_meth_rev : S.d_meth_fwd) : nullptr;
    Sink<WRITE> out{dst, 0ull};
    int idx = 0;
    for (int pi = 0; pi < m.length; ++pi) {
        const int p = m.start + pi;
        const int n = events_at(S, s_pos, s_extra, p);
        if (n == 0) continue;
        uint8_t kb[KM], mk[KM];
#pragma unroll
        for (int c = 0; c < KM; ++c) kb[c] = g[p + c];
        int kidx = 0;
#pragma unroll
        for (int c = 0; c < KM; ++c) {
            uint8_t ch;
            if (m.rev) {
                const uint8_t s = kb[KM - 1 - c];
                ch = s == 'A' ? 'T' : s == 'C' ? 'G' : s == 'G' ? 'C' : 'A';
            } else ch = kb[c];
            mk[c] = ch;
            kidx = kidx * 4 + base_code(ch);
        }
        const int mm = S.d_model_mean[kidx], ms = S.d_model_sd[kidx];
        int off = 0;
        if (mbm) {
            const long long gg = S.d_gbase[m.ci] + p;
#pragma unroll
            for (int c = 0; c < KM; ++c)
                if ((mbm[(gg + c) >> 5] >> ((gg + c) & 31)) & 1u) off += c_meth_off[c];
        }
        for (int j = 0; j < n; ++j, ++idx) {
            const unsigned long long v = H(s_ev, (unsigned long long)p, (unsigned long long)j);
            const unsigned long long v2 = H(s_ev2, (unsigned long long)p, (unsigned long long)j);
            const int e = m.e0 + (m.rev ? total - 1 - idx : idx);
            const int ssum = (int)((v >> 16) & 0xFFFull) + (int)((v >> 28) & 0xFFFull) + (int)((v >> 40) & 0xFFFull) +
                             (int)((v >> 52) & 0xFFFull) - 8190;
            int noise = (abs(ssum) * 240) / 2365;
            if (ssum < 0) noise = -noise;
            const int evc = mm + noise + off;
            const unsigned stdv = 500u + (unsigned)(v2 % 2500ull);
            const unsigned dur = 100u + (unsigned)((v2 >> 16) % 900ull);
            for (int c = 0; c < cname_len; ++c) out.put(cname[c]);
            out.put('\t');
            out.put_uint((unsigned)p);
            out.put('\t');
#pragma unroll
            for (int c = 0; c < KM; ++c) out.put(kb[c]);
            out.put('\t');
            for (int c = 0; c < nlen; ++c) out.put((uint8_t)name[c]);
            out.put('\t'); out.put('t'); out.put('\t');
            out.put_uint((unsigned)e);
            out.put('\t');
            out.put_centi(evc);
            out.put('\t');
            out.put_uint(stdv / 1000u); out.put('.'); out.put_fixed(stdv % 1000u, 3);
            out.put('\t');
            out.put('0'); out.put('.'); out.put_fixed(dur, 5);
            out.put('\t');
            if ((int)(v & 0xFFFFull) < S.p_nnn) {
                const char *t = "NNNNNN\t0.00\t0.00\tinf";
                for (int c = 0; t[c]; ++c) out.put((uint8_t)t[c]);
            } else {
#pragma unroll
                for (int c = 0; c < KM; ++c) out.put(mk[c]);
                out.put('\t');
                out.put_centi(mm);
                out.put('\t');
                out.put_centi(ms);
                out.put('\t');
                int z = (abs(evc - mm) * 100) / ms;
                if (evc < mm) z = -z;
                out.put_centi(z);
            }
            out.put('\n');
        }
    }
    return out.n;
}

__global__ void __launch_bounds__(128) k_synth_sizes(mc_synth_spec S, long long read0, long long n, unsigned long long *sizes) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    sizes[t] = gen_read<false>(S, read0 + t, nullptr);
}
__global__ void __launch_bounds__(128) k_synth_write(mc_synth_spec S, long long read0, long long n, const unsigned long long *offsets,
                                                    uint8_t *text) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    gen_read<true>(S, read0 + t, text + offsets[t]);
}
__global__ void __launch_bounds__(256) k_synth_genome(mc_synth_spec S, uint8_t *out, long long total) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    uint8_t ch = 'N';
    for (int ci = 0; ci < S.n_contigs; ++ci) {
        const long long b = S.d_gbase[ci];
        if (g >= b && g < b + S.d_contig_len[ci]) {
            const unsigned long long h = H(stream_seed(S.seed, S_REF), (unsigned long long)ci, (unsigned long long)(g - b));
            ch = (uint8_t)("ACGT"[(h >> 13) & 3ull]);
            break;
        }
    }
    out[g] = ch;
}

}  // namespace

extern "C" int mc_synth_genome(const mc_synth_spec *spec, uint8_t *d_genome_out, int64_t total_bits, void *stream) {
    MC_REQUIRE(spec && d_genome_out && total_bits > 0, "bad argument");
    k_synth_genome<<<(unsigned)((total_bits + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*spec, d_genome_out, total_bits);
    MC_LAUNCH_CHECK();
    return MC_OK;
}
extern "C" int mc_synth_sizes(const mc_synth_spec *spec, int64_t read0, int64_t n, int64_t n_total_reads, uint64_t *d_sizes, void *stream) {
    MC_REQUIRE(spec && d_sizes && n >= 0 && read0 >= 0 && read0 + n <= n_total_reads, "bad argument");
    if (n == 0) return MC_OK;
    k_synth_sizes<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*spec, read0, n, reinterpret_cast<unsigned long long *>(d_sizes));
    MC_LAUNCH_CHECK();
    return MC_OK;
}
extern "C" int mc_synth_write(const mc_synth_spec *spec, int64_t read0, int64_t n, int64_t n_total_reads, const uint64_t *d_offsets,
                              uint8_t *d_text, void *stream) {
    MC_REQUIRE(spec && d_offsets && d_text && n >= 0 && read0 >= 0 && read0 + n <= n_total_reads, "bad argument");
    if (n == 0) return MC_OK;
    k_synth_write<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*spec, read0, n,
                                                                              reinterpret_cast<const unsigned long long *>(d_offsets), d_text);
    MC_LAUNCH_CHECK();
    return MC_OK;
}
