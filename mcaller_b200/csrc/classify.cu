// Stage 6 (K3): per-observation classifier -- model[key].predict_proba([x])[0][1] and the 0.5 label threshold of
// the reference (extract_contexts.py:195-207) for the estimator types mCaller can be run with (-c NN/RF/LR/NBC;
// the estimator type is whatever the pickle holds, mCaller.py:137).
//
// Arithmetic is float64 like scikit-learn's.  The work per call is ~1.6 kFLOP (MLP 7-100-1) and there is one call
// per ~85 KB of TSV, so this stage is <1 % of the step; the layer widths (K=7, N=100) would leave a tcgen05 tile
// >90 % padding and bf16/tf32 operands cannot meet the 1e-5 probability tolerance, so the MLP runs on the FP64
// FMA pipe: one warp per call, lanes over hidden units, weights served from L1/L2 (7 KB per model).
#include "common.cuh"

namespace {

constexpr int MAX_WIDTH = 512;     // widest MLP layer supported
constexpr int WARPS = 8;

__device__ __forceinline__ double expit(double x) { return x < 0.0 ? exp(x) / (1.0 + exp(x)) : 1.0 / (1.0 + exp(-x)); }

__device__ __forceinline__ double activate(double v, int act) {
    switch (act) {
        case MC_ACT_TANH: return tanh(v);
        case MC_ACT_LOGISTIC: return expit(v);
        case MC_ACT_RELU: return v > 0.0 ? v : 0.0;
        default: return v;
    }
}

// sklearn MLPClassifier._forward_pass_fast (neural_network/_multilayer_perceptron.py) for a binary classifier:
// hidden activations, logistic output, P(class 1) = expit(z).
// Persistent warps: a block stages the weights and intercepts of both models in shared memory once (7-100-1: 7.2 KB per
// model) and its warps then stride over the calls, so the inner loops read nothing but shared memory; the features of a
// warp's next call are fetched while the current one is evaluated.  STAGED = false (models too wide for shared memory)
// reads the weights through the read-only cache instead.
template <bool STAGED>
__global__ void __launch_bounds__(WARPS * 32)
k_mlp(mc_call *__restrict__ calls, int64_t n, mc_model m0, mc_model m1, int width, int nw0, int nb0, int nw1, int nb1, int alias) {
    extern __shared__ __align__(16) double s_mlp[];          // [weights 0][biases 0][weights 1][biases 1][WARPS][2][width]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *s_act = s_mlp;
    if (STAGED) {
        for (int j = threadIdx.x; j < nw0; j += WARPS * 32) s_mlp[j] = __ldg(m0.d_weights + j);
        for (int j = threadIdx.x; j < nb0; j += WARPS * 32) s_mlp[nw0 + j] = __ldg(m0.d_biases + j);
        for (int j = threadIdx.x; j < nw1; j += WARPS * 32) s_mlp[nw0 + nb0 + j] = __ldg(m1.d_weights + j);
        for (int j = threadIdx.x; j < nb1; j += WARPS * 32) s_mlp[nw0 + nb0 + nw1 + j] = __ldg(m1.d_biases + j);
        s_act = s_mlp + nw0 + nb0 + nw1 + nb1;
        __syncthreads();
    }
    double *a = s_act + (size_t)warp * 2 * width, *b = a + width;
    const int64_t stride = (int64_t)gridDim.x * WARPS;
    int64_t i = (int64_t)blockIdx.x * WARPS + warp;
    // features (and row kind / model selector) of the call the warp handles next
    auto fetch = [&](int64_t idx, double &f, int &kind_sel) {
        f = 0.0;
        kind_sel = -1;
        if (idx < n) {
            const mc_call &c = calls[idx];
            kind_sel = c.kind == MC_CALL ? (int)c.model_sel : -1;
            if (lane <= MC_MAXK) f = c.feat[lane];
        }
    };
    double f_cur;
    int ks_cur;
    fetch(i, f_cur, ks_cur);
    for (; i < n; i += stride) {
        double f_next;
        int ks_next;
        fetch(i + stride, f_next, ks_next);
        if (ks_cur >= 0) {
            const mc_model &m = ks_cur ? m1 : m0;
            // model 1 follows model 0 in shared memory unless both selectors name the same model (alias)
            const double *w = STAGED ? ((ks_cur && !alias) ? s_mlp + nw0 + nb0 : s_mlp) : m.d_weights;
            const double *bi = STAGED ? ((ks_cur && !alias) ? s_mlp + nw0 + nb0 + nw1 : s_mlp + nw0) : m.d_biases;
            auto ld = [&](const double *p) { return STAGED ? *p : __ldg(p); };
            double *x = a, *y = b;
            if (lane < m.sizes[0]) x[lane] = f_cur;
            __syncwarp();
            for (int l = 0; l < m.n_layers; ++l) {
                const int ni = m.sizes[l], no = m.sizes[l + 1];
                const bool last = (l + 1 == m.n_layers);
                if (no >= 8) {
                    // four outputs per lane at a time: independent FMA / activation chains interleave and hide each other's latency
                    for (int o0 = 0; o0 < no; o0 += 128) {
                        double acc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = 0.0;
                        for (int k = 0; k < ni; ++k) {
                            const double ak = x[k];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int o = o0 + lane + 32 * j;
                                if (o < no) acc[j] = fma(ak, ld(w + (size_t)k * no + o), acc[j]);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int o = o0 + lane + 32 * j;
                            if (o < no) {
                                const double z = acc[j] + ld(bi + o);          // dot product first, then the intercept (as sklearn)
                                y[o] = last ? z : activate(z, m.hidden_act);
                            }
                        }
                    }
                } else {
                    for (int o = 0; o < no; ++o) {
                        double acc = 0.0;
                        for (int k = lane; k < ni; k += 32) acc = fma(x[k], ld(w + (size_t)k * no + o), acc);
                        for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
                        acc += ld(bi + o);
                        if (lane == 0) y[o] = last ? acc : activate(acc, m.hidden_act);
                    }
                }
                __syncwarp();
                double *t = x; x = y; y = t;
                w += (size_t)ni * no;
                bi += no;
            }
            const double out = expit(x[0]);
            __syncwarp();
            if (lane == 0) {
                mc_call &c = calls[i];
                c.prob = out;
                c.label = (uint8_t)(out >= 0.5);
            }
        }
        f_cur = f_next;
        ks_cur = ks_next;
    }
}

// The shape every shipped pickle has -- one hidden layer, at most 9 inputs (k + 1 <= 9), one logistic output -- four lanes
// per call (eight calls per warp): a lane keeps its call's features in registers and owns every fourth hidden unit, so the
// activation (float64 tanh: ~80 % of the instructions) is evaluated 25 times per lane for 8 calls instead of 4 times for
// one, nothing passes through shared memory but the staged weights, and the output dot product is finished with two
// shuffles.  Same arithmetic as k_mlp: dot product in input order, then the intercept, activation, logistic output.
constexpr int MLP_MAX_IN = MC_MAXK + 1;
__global__ void __launch_bounds__(WARPS * 32)
k_mlp_1hidden(mc_call *__restrict__ calls, int64_t n, mc_model m0, mc_model m1, int nw0, int nb0, int nw1, int nb1, int alias) {
    extern __shared__ __align__(16) double s_mlp[];          // [weights 0][biases 0][weights 1][biases 1]
    for (int j = threadIdx.x; j < nw0; j += WARPS * 32) s_mlp[j] = __ldg(m0.d_weights + j);
    for (int j = threadIdx.x; j < nb0; j += WARPS * 32) s_mlp[nw0 + j] = __ldg(m0.d_biases + j);
    for (int j = threadIdx.x; j < nw1; j += WARPS * 32) s_mlp[nw0 + nb0 + j] = __ldg(m1.d_weights + j);
    for (int j = threadIdx.x; j < nb1; j += WARPS * 32) s_mlp[nw0 + nb0 + nw1 + j] = __ldg(m1.d_biases + j);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane & 3, grp = lane >> 2;
    const int64_t stride = (int64_t)gridDim.x * WARPS * 8;
    for (int64_t base = ((int64_t)blockIdx.x * WARPS + warp) * 8; base < n; base += stride) {      // warp-uniform
        const int64_t i = base + grp;
        int ks = -1;
        double x[MLP_MAX_IN];
#pragma unroll
        for (int k = 0; k < MLP_MAX_IN; ++k) x[k] = 0.0;
        if (i < n) {
            const mc_call &c = calls[i];
            if (c.kind == MC_CALL) {
                ks = (int)c.model_sel;
#pragma unroll
                for (int k = 0; k < MLP_MAX_IN; ++k) x[k] = c.feat[k];
            }
        }
        const bool second = ks > 0;
        const mc_model &m = second ? m1 : m0;
        const int ni = m.sizes[0], H = m.sizes[1], act = m.hidden_act;
        const double *w0 = (second && !alias) ? s_mlp + nw0 + nb0 : s_mlp;
        const double *b0 = (second && !alias) ? s_mlp + nw0 + nb0 + nw1 : s_mlp + nw0;
        const double *w1 = w0 + (size_t)ni * H, *b1 = b0 + H;
        double out = 0.0;
        if (ks >= 0) {
            for (int o0 = g; o0 < H; o0 += 16) {              // four hidden units per lane at a time: independent chains
                double z[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) z[j] = 0.0;
#pragma unroll
                for (int k = 0; k < MLP_MAX_IN; ++k) {
                    if (k < ni) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int o = o0 + 4 * j;
                            if (o < H) z[j] = fma(x[k], w0[(size_t)k * H + o], z[j]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = o0 + 4 * j;
                    if (o < H) out = fma(activate(z[j] + b0[o], act), w1[o], out);     // dot product first, then the intercept
                }
            }
        }
        out += __shfl_xor_sync(0xffffffffu, out, 1);
        out += __shfl_xor_sync(0xffffffffu, out, 2);
        if (ks >= 0 && g == 0) {
            const double p = expit(out + b1[0]);
            mc_call &c = calls[i];
            c.prob = p;
            c.label = (uint8_t)(p >= 0.5);
        }
    }
}

// LogisticRegression.predict_proba (binary): expit(x.w + b); GaussianNB.predict_proba: exp(jll_1 - logsumexp(jll))
__global__ void __launch_bounds__(256) k_linear(mc_call *__restrict__ calls, int64_t n, mc_model m0, mc_model m1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mc_call &c = calls[i];
    if (c.kind != MC_CALL) return;
    const mc_model &m = c.model_sel ? m1 : m0;
    double p;
    if (m.kind == MC_LR) {
        double z = __ldg(m.d_biases);
        for (int k = 0; k < m.n_in; ++k) z += c.feat[k] * __ldg(m.d_weights + k);
        p = expit(z);
    } else {
        const double *theta = m.d_weights, *var = m.d_weights + 2 * m.n_in;
        double jll[2];
        for (int cl = 0; cl < 2; ++cl) {
            double nij = 0.0, s = 0.0;
            for (int k = 0; k < m.n_in; ++k) nij += log(2.0 * 3.14159265358979323846 * __ldg(var + cl * m.n_in + k));
            nij *= -0.5;
            for (int k = 0; k < m.n_in; ++k) {
                const double d = c.feat[k] - __ldg(theta + cl * m.n_in + k);
                s += d * d / __ldg(var + cl * m.n_in + k);
            }
            jll[cl] = __ldg(m.d_biases + cl) + nij - 0.5 * s;
        }
        const double mx = jll[0] > jll[1] ? jll[0] : jll[1];
        const double lse = mx + log(exp(jll[0] - mx) + exp(jll[1] - mx));
        p = exp(jll[1] - lse);
    }
    c.prob = p;
    c.label = (uint8_t)(p >= 0.5);
}

// RandomForestClassifier.predict_proba: mean over trees of the leaf class-1 fraction; the walk compares float32(x)
// with the float64 threshold (sklearn tree/_tree.pyx).  A block owns 256 calls and streams the trees through shared
// memory (node table of one tree at a time), so every node read in the walk is an smem read.
__global__ void __launch_bounds__(256)
k_rf(mc_call *__restrict__ calls, int64_t n, mc_model m0, mc_model m1, int max_nodes) {
    extern __shared__ __align__(16) uint8_t rf_smem[];
    double *s_thr = reinterpret_cast<double *>(rf_smem);
    double *s_p1 = s_thr + max_nodes;
    int32_t *s_left = reinterpret_cast<int32_t *>(s_p1 + max_nodes);
    int32_t *s_right = s_left + max_nodes;
    int32_t *s_feat = s_right + max_nodes;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n && calls[i].kind == MC_CALL;
    float x[MC_MAXK + 1];
    int sel = 0;
    if (active) {
        sel = calls[i].model_sel;
        for (int k = 0; k <= MC_MAXK; ++k) x[k] = (float)calls[i].feat[k];
    }
    double acc = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        const mc_model &m = pass ? m1 : m0;
        // skip a model nobody in the block uses
        const int any = __syncthreads_or(active && sel == pass);
        if (!any || m.n_trees == 0) continue;
        for (int t = 0; t < m.n_trees; ++t) {
            const int o0 = __ldg(m.d_tree_off + t), nn = __ldg(m.d_tree_off + t + 1) - o0;
            __syncthreads();
            for (int j = threadIdx.x; j < nn; j += blockDim.x) {
                s_thr[j] = __ldg(m.d_threshold + o0 + j);
                s_p1[j] = __ldg(m.d_leaf_p1 + o0 + j);
                s_left[j] = __ldg(m.d_left + o0 + j);
                s_right[j] = __ldg(m.d_right + o0 + j);
                s_feat[j] = __ldg(m.d_feature + o0 + j);
            }
            __syncthreads();
            if (active && sel == pass) {
                int node = 0;
                while (s_left[node] >= 0) node = ((double)x[s_feat[node]] <= s_thr[node]) ? s_left[node] : s_right[node];
                acc += s_p1[node];
            }
        }
    }
    if (active) {
        const mc_model &m = sel ? m1 : m0;
        const double p = acc / (double)m.n_trees;
        calls[i].prob = p;
        calls[i].label = (uint8_t)(p >= 0.5);
    }
}

// ---- stage 7 (K4): per-site histogram (make_bed.py:86-96) ----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_hist(const mc_call *__restrict__ calls, int64_t n, uint32_t *__restrict__ depth, uint32_t *__restrict__ meth,
       unsigned long long *__restrict__ first, int64_t n_sites, unsigned long long row_base,
       unsigned long long *__restrict__ d_skipped) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const mc_call &c = calls[i];
    if (c.kind != MC_CALL) return;
    if (c.close_rec == 0xFFFFFFFFu || c.chrom_contig != c.win_contig || c.site < 0 || c.site >= n_sites || c.err) {
        atomicAdd(d_skipped, 1ull);
        return;
    }
    atomicAdd(depth + c.site, 1u);
    if (c.label) atomicAdd(meth + c.site, 1u);
    atomicMin(first + c.site, row_base + (unsigned long long)i);
}

// row statistics without a D2H copy of the rows: [0] calls closed in this chunk, [1] calls still pending,
// [2] too-many-skips events closed in this chunk, [3] multi-M events, [4] rows with an error flag, [5] calls labelled
// methylated, [6] too-many-skips events still pending (a window open at the end of the chunk is closed -- and only then
// counted, extract_contexts.py:179/:238 -- by the next kept line of the file; the last one of a file never is)
__global__ void __launch_bounds__(256) k_count_calls(const mc_call *__restrict__ calls, int64_t n, unsigned long long *__restrict__ out) {
    __shared__ unsigned int s[7];
    if (threadIdx.x < 7) s[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const mc_call &c = calls[i];
        if (c.kind == MC_CALL) {
            atomicAdd(&s[c.close_rec == 0xFFFFFFFFu ? 1 : 0], 1u);
            if (c.label) atomicAdd(&s[5], 1u);
        } else if (c.kind == MC_TOO_MANY_SKIPS) atomicAdd(&s[c.close_rec == 0xFFFFFFFFu ? 6 : 2], 1u);
        else atomicAdd(&s[3], 1u);
        if (c.err) atomicAdd(&s[4], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 7 && s[threadIdx.x]) atomicAdd(out + threadIdx.x, (unsigned long long)s[threadIdx.x]);
}

}  // namespace

extern "C" int mc_count_calls(const mc_call *d_calls, int64_t n_calls, uint64_t *d_out, void *stream) {
    MC_REQUIRE(d_calls && d_out, "null pointer");
    if (n_calls <= 0) return MC_OK;
    k_count_calls<<<(unsigned)((n_calls + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_calls, n_calls,
                                                                                       reinterpret_cast<unsigned long long *>(d_out));
    MC_LAUNCH_CHECK();
    return MC_OK;
}

extern "C" int mc_classify(mc_call *d_calls, int64_t n_calls, const mc_model *models, void *stream) {
    MC_REQUIRE(d_calls && models, "null pointer");
    if (n_calls <= 0) return MC_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const mc_model &m0 = models[0], &m1 = models[1];
    MC_REQUIRE(m0.n_in >= 1 && m0.n_in <= MC_MAXK + 1, "model input width out of range");
    switch (m0.kind) {
        case MC_MLP: {
            MC_REQUIRE(m0.n_layers >= 1 && m0.n_layers < 8, "MLP depth out of range");
            for (int l = 0; l <= m0.n_layers; ++l) MC_REQUIRE(m0.sizes[l] >= 1 && m0.sizes[l] <= MAX_WIDTH, "MLP layer too wide");
            MC_REQUIRE(m0.sizes[m0.n_layers] == 1, "MLP must have one logistic output");
            int width = 8;
            for (int l = 0; l <= m0.n_layers; ++l) width = m0.sizes[l] > width ? m0.sizes[l] : width;
            for (int l = 0; l <= m1.n_layers && m1.kind == MC_MLP; ++l) width = m1.sizes[l] > width ? m1.sizes[l] : width;
            width = (width + 3) & ~3;
            auto count = [](const mc_model &m, int &nw, int &nb) {
                nw = nb = 0;
                if (m.kind != MC_MLP) return;
                for (int l = 0; l < m.n_layers; ++l) { nw += m.sizes[l] * m.sizes[l + 1]; nb += m.sizes[l + 1]; }
            };
            int nw0, nb0, nw1, nb1;
            count(m0, nw0, nb0);
            count(m1, nw1, nb1);
            const int alias = (m1.kind == MC_MLP && m1.d_weights == m0.d_weights) ? 1 : 0;
            if (alias) nw1 = nb1 = 0;                             // one model given twice: stage it once
            int dev = 0, sms = 0;
            MC_CUDA_CHECK(cudaGetDevice(&dev));
            MC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            const size_t act_bytes = sizeof(double) * 2 * (size_t)width * WARPS;
            const size_t staged_bytes = act_bytes + sizeof(double) * (size_t)(nw0 + nb0 + nw1 + nb1);
            int64_t blocks = (n_calls + WARPS - 1) / WARPS;
            if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;           // persistent: warps stride over the calls
            auto one_hidden = [](const mc_model &m) {
                return m.kind == MC_MLP && m.n_layers == 2 && m.sizes[0] <= MLP_MAX_IN && m.sizes[2] == 1;
            };
            const size_t weight_bytes = sizeof(double) * (size_t)(nw0 + nb0 + nw1 + nb1);
            if (one_hidden(m0) && (alias || m1.kind != MC_MLP || one_hidden(m1)) && weight_bytes <= 100 * 1024) {
                int64_t b2 = (n_calls + WARPS * 8 - 1) / (WARPS * 8);
                if (b2 > (int64_t)sms * 8) b2 = (int64_t)sms * 8;
                MC_CUDA_CHECK(cudaFuncSetAttribute(k_mlp_1hidden, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)weight_bytes));
                k_mlp_1hidden<<<(unsigned)b2, WARPS * 32, weight_bytes, st>>>(d_calls, n_calls, m0, m1, nw0, nb0, nw1, nb1, alias);
            } else if (staged_bytes <= 100 * 1024) {
                MC_CUDA_CHECK(cudaFuncSetAttribute(k_mlp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)staged_bytes));
                k_mlp<true><<<(unsigned)blocks, WARPS * 32, staged_bytes, st>>>(d_calls, n_calls, m0, m1, width, nw0, nb0, nw1, nb1, alias);
            } else {
                k_mlp<false><<<(unsigned)blocks, WARPS * 32, act_bytes, st>>>(d_calls, n_calls, m0, m1, width, 0, 0, 0, 0, 0);
            }
            break;
        }
        case MC_LR:
        case MC_GNB:
            k_linear<<<(unsigned)((n_calls + 255) / 256), 256, 0, st>>>(d_calls, n_calls, m0, m1);
            break;
        case MC_RF: {
            int mx = m0.max_nodes > m1.max_nodes ? m0.max_nodes : m1.max_nodes;
            MC_REQUIRE(mx >= 1, "RF max_nodes missing");
            const size_t smem = (size_t)mx * (8 + 8 + 4 + 4 + 4);
            MC_REQUIRE(smem <= 200 * 1024, "RF tree too large for shared-memory staging");
            MC_CUDA_CHECK(cudaFuncSetAttribute(k_rf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_rf<<<(unsigned)((n_calls + 255) / 256), 256, smem, st>>>(d_calls, n_calls, m0, m1, mx);
            break;
        }
        default:
            MC_REQUIRE(false, "unknown model kind");
    }
    MC_LAUNCH_CHECK();
    return MC_OK;
}

extern "C" int mc_hist_accumulate(const mc_call *d_calls, int64_t n_calls, uint32_t *d_depth, uint32_t *d_meth, uint64_t *d_first,
                                  int64_t n_sites, uint64_t row_base, uint64_t *d_skipped, void *stream) {
    MC_REQUIRE(d_calls && d_depth && d_meth && d_first && d_skipped, "null pointer");
    if (n_calls <= 0) return MC_OK;
    k_hist<<<(unsigned)((n_calls + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_calls, n_calls, d_depth, d_meth, reinterpret_cast<unsigned long long *>(d_first), n_sites,
        (unsigned long long)row_base, reinterpret_cast<unsigned long long *>(d_skipped));
    MC_LAUNCH_CHECK();
    return MC_OK;
}
