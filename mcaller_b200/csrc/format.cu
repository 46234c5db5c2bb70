// Host-side writer of the `.diffs.<k>` rows (reference extract_contexts.py:216 + writefi :83-86), native so that runs with
// millions of calls are not bound by Python string formatting.  Floats are printed like str(np.float64) / repr(float):
// the shortest digit string that round-trips (std::to_chars), fixed notation for 1e-4 <= |x| < 1e16, else scientific
// with a two-digit exponent; empty columns print the integer 0; the probability is np.round(p, 2) (rint(p*100)/100).
#include <charconv>
#include <cmath>
#include <cstring>
#include "common.cuh"

namespace {

struct Out {
    char *p, *end;
    bool ok = true;
    void put(char c) { if (p < end) *p++ = c; else ok = false; }
    void put(const char *s, size_t n) { if ((size_t)(end - p) >= n) { memcpy(p, s, n); p += n; } else ok = false; }
    void put(const char *s) { put(s, strlen(s)); }
};

void put_repr(Out &o, double v) {
    if (std::isnan(v)) { o.put("nan"); return; }
    if (std::isinf(v)) { o.put(v < 0 ? "-inf" : "inf"); return; }
    if (v == 0.0) { o.put(std::signbit(v) ? "-0.0" : "0.0"); return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);   // shortest round-trip digits
    *r.ptr = 0;
    // parse  [-]d[.ddd]e[+-]XX
    const char *s = buf;
    if (*s == '-') { o.put('-'); ++s; }
    char digits[32];
    int nd = 0;
    digits[nd++] = *s++;
    if (*s == '.') { ++s; while (*s != 'e') digits[nd++] = *s++; }
    ++s;                                   // 'e'
    const int e10 = atoi(s);
    if (e10 >= -4 && e10 < 16) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) o.put(i < nd ? digits[i] : '0');
            o.put('.');
            if (nd > e10 + 1) o.put(digits + e10 + 1, (size_t)(nd - e10 - 1)); else o.put('0');
        } else {
            o.put("0.");
            for (int i = 0; i < -e10 - 1; ++i) o.put('0');
            o.put(digits, (size_t)nd);
        }
    } else {
        o.put(digits[0]);
        if (nd > 1) { o.put('.'); o.put(digits + 1, (size_t)(nd - 1)); }
        o.put('e');
        o.put(e10 < 0 ? '-' : '+');
        const int a = e10 < 0 ? -e10 : e10;
        char eb[8];
        const int n = snprintf(eb, sizeof(eb), "%02d", a);
        o.put(eb, (size_t)n);
    }
}

void put_int(Out &o, long long v) {
    char b[32];
    auto r = std::to_chars(b, b + sizeof(b), v);
    o.put(b, (size_t)(r.ptr - b));
}

char comp(char c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'N': return 'N'; case 'M': return 'M'; }
    return 0;
}

}  // namespace

extern "C" int64_t mc_format_rows(const mc_call *h_calls, int64_t n_calls, const uint8_t *h_text, const char *const *contig_names,
                                  const char *const *marked_fwd, const char *const *marked_rev, const int64_t *contig_len,
                                  int32_t n_contigs, int32_t k, const char *base_label, const char *mod_label, int32_t with_prob,
                                  char *out, int64_t out_cap) {
    if (!h_calls || !h_text || !contig_names || !marked_fwd || !marked_rev || !contig_len || !out || k < 1 || k > MC_MAXK) {
        mc_set_error("mc_format_rows: bad argument");
        return MC_EINVAL;
    }
    Out o{out, out + out_cap};
    for (int64_t i = 0; i < n_calls; ++i) {
        const mc_call &c = h_calls[i];
        if (c.kind != MC_CALL || c.close_rec == 0xFFFFFFFFu) continue;
        if (c.err) { mc_set_error("row %lld (position %d) carries error flags 0x%x", (long long)i, c.mpos, (unsigned)c.err); return -100 - (int64_t)c.err; }
        if (c.chrom_contig >= n_contigs || c.win_contig >= n_contigs) { mc_set_error("mc_format_rows: contig index out of range"); return MC_EINVAL; }
        o.put(contig_names[c.chrom_contig]);
        o.put('\t');
        o.put(reinterpret_cast<const char *>(h_text) + c.read_off, (size_t)c.read_len);
        o.put('\t');
        put_int(o, c.mpos);
        o.put('\t');
        // context = revcomp(last_ref[mpos-k+1 : mpos+k], last_rev)   (extract_contexts.py:194)
        const char *src = c.rev ? marked_rev[c.win_contig] : marked_fwd[c.win_contig];
        const int64_t a = (int64_t)c.mpos - k + 1, b = (int64_t)c.mpos + k;
        if (a < 0 || b > contig_len[c.win_contig]) { mc_set_error("mc_format_rows: context out of range"); return MC_EINVAL; }
        for (int64_t j = 0; j < b - a; ++j) o.put(c.rev ? comp(src[b - 1 - j]) : src[a + j]);
        o.put('\t');
        for (int j = 0; j <= k; ++j) {
            if (j) o.put(',');
            if (j < k && ((c.empty_mask >> j) & 1u)) o.put('0'); else put_repr(o, c.feat[j]);
        }
        o.put('\t');
        o.put(c.rev ? '-' : '+');
        if (with_prob) {
            o.put('\t');
            o.put(c.prob >= 0.5 ? mod_label : base_label);
            o.put('\t');
            put_repr(o, std::rint(c.prob * 100.0) / 100.0);
        }
        o.put('\n');
        if (!o.ok) { mc_set_error("mc_format_rows: output buffer too small"); return MC_ECAPACITY; }
    }
    return (int64_t)(o.p - out);
}
