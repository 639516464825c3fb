// libgcpb200: context, weight packing and the rollout schedule behind the C ABI in include/gcpb200.h.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <vector>

// GCPB200_VERIFY (tests/cuda/libgcpb200_verify.so only, never the shipped library): compiles the SIMT verification
// kernels (gemm_ref_kernel, dec_tail_ref_kernel, dec_tail_nll_kernel) and the GCPB200_* environment switches used for
// A/B measurements.  Without it the library has one code path and reads no environment variable.
#ifdef GCPB200_VERIFY
#define GCP_VERIFY 1
#else
#define GCP_VERIFY 0
#endif

#include "../../include/gcpb200.h"
#include "dec_tail3.cuh"
#include "dtw_kernels.cuh"
#include "gemm_host.cuh"
#include "kernels_misc.cuh"
#include "mlp_fused.cuh"
#include "train_kernels.cuh"

using namespace gcp;

// ---------------------------------------------------------------------------------------------
// error reporting
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
extern "C" void gcp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* gcpb200_last_error(void) { return g_err; }
extern "C" const char* gcpb200_version(void) { return "gcpb200 0.1.0 (sm_100a)"; }

#define CHECK(x)             \
    do {                     \
        if ((x) != 0) return -1; \
    } while (0)

// model dimensions of the 25-room GCP-tree (experiments/control/25room/gcp_tree/mod_hyper.py:33-54)
static const int DEPTH = 8, N_NODES = 255;        // 25-room defaults (and the maxima); the context carries the live values
static const int NZ_ENC = 128, NZ_VAE = 256, NZ_MID = 128, HID = 512, N_LSTM = 3, STATE = 3072;
static const int MAX_LEN = 200, INIT_MID = 32;
static const int REFIT_SPLITS = 16;

// ---------------------------------------------------------------------------------------------
// device containers
// ---------------------------------------------------------------------------------------------
struct DevMat {           // packed weight matrix [N][K] bf16 (K-major) + fp32 bias in packed column order
    bf16* w = nullptr;
    float* bias = nullptr;
    int N = 0, K = 0;
    CUtensorMap map_box[5];   // TMA boxes of 16, 32, 64, 128, 256 weight rows (BN / cluster size)
};
struct DevBuf {           // bf16 activation array [rows][ld] with a TMA map over its full extent
    bf16* p = nullptr;
    size_t rows = 0;
    int ld = 0;
    CUtensorMap map;
};
struct Mlp {              // BaseProcessingNet: in(+bias,LReLU) -> 3x[lin, GroupNorm(8), LReLU] -> head(+bias)
    DevMat in, mid[3], head;
    float* gam[3] = {nullptr, nullptr, nullptr};
    float* bet[3] = {nullptr, nullptr, nullptr};
    int gn_group = 16;    // channels per group
    int mid_k = 128;      // K of the mid/head layers (mid width padded to 64)
    int mid_valid = 128;  // columns written by in/mid layers
    int n_out = 0;
    int n_mid = 3;        // number of [lin, GroupNorm, LReLU] layers (odd, so the body ends in ctx->tb)
};
struct SeqW {             // VRNNCell of the sequential GCP model (prior + gen_lstm; blox/torch/models/vrnn.py:24-52)
    Mlp prior;            // 128 -> 128 x3 -> 512 (mu | log sigma, interleaved for the reparametrisation epilogue)
    Mlp init;             // init_module: 256 -> 128 -> 6144, head rows packed [h0 h1 h2 | c0 c1 c2]
    DevMat embed_main, embed_ctx, lstm[3], out;
};
struct LevelW {
    Mlp prior;
    Mlp q;                // approximate posterior q(z | e_l, e_r, e_tilde) (training path only; tree/inference.py:16-36)
    Mlp init;             // level 0 only; head rows [0,3072) -> left state, [3072,6144) -> right state
    DevMat init_head_r;   // second half of the init head
    DevMat proj, embed_main, lstm[3], out;
};

struct gcpb200_ctx {
    gcpb200_config cfg;
    int Bp_max = 0, sms = 0, slot_chunk = 64, max_cluster = 2;
#if GCP_VERIFY
    bool use_ref = false;                     // verification build only: SIMT kernels instead of tcgen05
#else
    static constexpr bool use_ref = false;    // the shipped library has exactly one code path
#endif
    bool weights_loaded = false;
    int64_t launches = 0;
    std::vector<void*> allocs;
    // weights
    LevelW lvl[8];
    SeqW seqw;
    int model = 0, n_slots = 257, lstm_hid = 512;
    // tree shape (gcpb200_config.hierarchy_levels / max_seq_len / tied_layers): 25-room = 8 levels, 255 nodes, 200 frames, one
    // TreeModule per level; 9-room = 7 levels, 127 nodes, 100 frames, ONE TreeModule for all levels
    int depth = 8, n_nodes = 255, max_len = 200;
    bool tied = false;
    DevBuf hs[2];            // sequential: bf16 h state [Bp][3 * 1024], ping-pong by step parity
    float* cs = nullptr;     // sequential: fp32 c state [Bp][3 * 1024]
    // sequential: the 199-step recurrence (~2000 dependent launches) is captured once per (noise buffer, prior output
    // buffers, B) into a CUDA graph and replayed; the host cost of a rollout drops from ~35 ms of launches to one
    struct SeqGraph { const void *z, *mu, *ls; int B; cudaGraphExec_t exec; int64_t launches; };
    std::vector<SeqGraph> seq_graphs;
    bool seq_graph_on = true, seq_warm = false;
    cudaStream_t seq_stream = nullptr;   // capture / replay stream (the caller's stream may be the legacy default stream,
    cudaEvent_t ev_seq_fork = nullptr, ev_seq_join = nullptr;   // which cannot be captured); fork/join by events
    Mlp length_pred, existence, inv_mdl, state_reg, cost_mdl, distance_pred;
    bool has_cost = false, has_inv = false, has_state = false;
    int pair_rows = 0;       // rows of the `pairs` / `rowcost` scratch arrays
    DevBuf seqb;             // bf16 copy of the pruned latent sequences, [cand][MAX_LEN + 1][128] (last row zero)
    EncoderWeights enc;
    const float *enc_w1t = nullptr, *enc_w2t = nullptr, *enc_w3t = nullptr;   // k-major copies for the batch-stat encoder
    DevMat dec1, dec2x, dec2s, dec3;
    bf16 *w4 = nullptr, *w5 = nullptr, *w4p = nullptr, *w5p = nullptr, *z4 = nullptr, *z5 = nullptr, *s4 = nullptr;
    float *b4 = nullptr, *b5 = nullptr, *b5h = nullptr;
    bf16 *z5m = nullptr, *z5s = nullptr;      // training NLL: unscaled Toeplitz arrays of the 15 mean / 15 log-scale channels
    float *b5m = nullptr, *b5s = nullptr;
    // workspace
    float* lat_f32 = nullptr;
    DevBuf lat, hid, xa, xb, zeta, sh, ta, tb, s2b, x1, x2, x3, pairs;
    float *ctxb = nullptr, *logits = nullptr, *s0 = nullptr, *s2 = nullptr, *rowbias2 = nullptr;
    bf16* skip_up = nullptr;
    float *exist_slot = nullptr, *e_df = nullptr, *seq = nullptr, *rowcost = nullptr, *goal_tail = nullptr;
    long long* end_ind = nullptr;
    long long* scratch_ei = nullptr;
    long long* scratch_given = nullptr;
    int* frame_node = nullptr;
    // planner mode "decode kept nodes only": compacted latent rows + their (candidate, node), per-image L2 sums
    DevBuf latc;
    int *row_cand = nullptr, *row_node = nullptr, *row_off = nullptr, *n_rows = nullptr;
    float* frame_sq = nullptr;
    int *tree_tiles = nullptr, *tree_rows = nullptr;   // per-level work lists of the pruned recursion (tree_worklists_kernel)
    unsigned long long* topk_sel = nullptr;   // elite selection: the k selected composite keys (grown on demand)
    int topk_cap = 0;
    double* refit_part = nullptr;             // refit: per-split (sum, sum of squares) [REFIT_SPLITS][255*256][2]
    // overlapped upload of host noise (gcpb200_rollout_io.z_host)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy_start = nullptr, ev_copy[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // side stream of the tree rollout: what depends on the encoder / rollout length only (frame map, kept-row offsets, the
    // decoder's per-call constants) runs beside the tree recursion instead of between it and the decoder
    cudaStream_t prep_stream = nullptr;
    cudaEvent_t ev_prep_fork = nullptr, ev_prep_join = nullptr;
    // second side stream: the parent-state projection GEMM of a tree level that does not fill the GPU runs beside that
    // level's prior -> reparametrisation -> embed chain (it needs the previous level's LSTM state only)
    cudaStream_t proj_stream = nullptr;
    cudaEvent_t ev_proj_fork = nullptr, ev_proj_join = nullptr;
    // ---- training-phase forward + loss (gcpb200_forward_loss)
    bool has_train = false;
    DevMat dec1t, dec2xt, dec2st, dec3t;          // decoder layers 1-3 without BatchNorm folding
    float *bn_g[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // BatchNorm affine: encoder pyramid-0 (32), pyramid-1 (64),
    float *bn_b[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // decoder net (64), pyramid-1 (32), pyramid-0 (16)
    float *ie_w[3] = {nullptr, nullptr, nullptr}, *ie_b[3] = {nullptr, nullptr, nullptr};   // inf_encoder conv1d, [k][cin][128]
    float *ie_gn_g = nullptr, *ie_gn_b = nullptr;
    struct TrainWS {
        int Bcap = 0;
        float *y1 = nullptr, *y2 = nullptr, *enc_seq = nullptr, *h1 = nullptr, *h2 = nullptr, *inf_seq = nullptr;
        double* stats = nullptr;                  // one zeroed block: st1 [3][32][2] | st2 [3][64][2] | gn [128][8][2] | d1 [64][2] | d2 | d3
        int* tstep = nullptr;
        unsigned char* keep = nullptr;
        DevBuf x1, x2, x3;                        // raw / normalised decoder activations of all 255 nodes
        float *pq[4] = {nullptr, nullptr, nullptr, nullptr};   // p_mu, p_ls, q_mu, q_ls [B][255][256]
        float *raw_mu = nullptr, *raw_ls = nullptr;   // raw head outputs of one 64-node chunk [128][64][1024][16]
        float *nll_bt = nullptr, *kl_b = nullptr, *reg = nullptr, *inv_pred = nullptr, *cost_pred = nullptr, *exist_df = nullptr, *cost_tgt = nullptr;
    } tw;
    // optional phase profiling (CUDA events on the caller's stream)
    bool profile = false;
    struct ProfSpan { int phase; cudaEvent_t a, b; };
    std::vector<ProfSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[GCPB200_N_PHASES] = {0, 0, 0, 0, 0, 0};
    long long prof_tail_images = 0, prof_tail_launches = 0;
};

static cudaEvent_t prof_event(gcpb200_ctx* c) {
    cudaEvent_t e;
    if (!c->ev_pool.empty()) {
        e = c->ev_pool.back();
        c->ev_pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}
struct ProfScope {
    gcpb200_ctx* c;
    cudaStream_t st;
    size_t idx = 0;
    bool on;
    ProfScope(gcpb200_ctx* c_, cudaStream_t st_, int phase) : c(c_), st(st_), on(c_->profile) {
        if (!on) return;
        gcpb200_ctx::ProfSpan sp{phase, prof_event(c), prof_event(c)};
        cudaEventRecord(sp.a, st);
        idx = c->spans.size();
        c->spans.push_back(sp);
    }
    ~ProfScope() {
        if (on) cudaEventRecord(c->spans[idx].b, st);
    }
};

// GCPB200_TRACE=1: per-call timeline of the rollout on stderr (events on the rollout and the copy stream, relative to
// the start of the call).  Diagnostic only: it synchronises the device at the end of every traced call.
struct TraceMark {
    const char* name;
    cudaEvent_t ev;
};
static std::vector<TraceMark>& trace_marks() {
    static std::vector<TraceMark> v;
    return v;
}
static bool trace_on() {
#if GCP_VERIFY
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GCPB200_TRACE");
        on = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
#else
    return false;
#endif
}
static void trace_mark(cudaStream_t st, const char* name) {
    if (!trace_on()) return;
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    trace_marks().push_back({name, ev});
}
static void trace_dump() {
    if (!trace_on() || trace_marks().empty()) return;
    cudaDeviceSynchronize();
    auto& v = trace_marks();
    fprintf(stderr, "[gcpb200 trace]");
    for (size_t i = 0; i < v.size(); ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, v[0].ev, v[i].ev);
        fprintf(stderr, " %s=%.2f", v[i].name, t);
    }
    fprintf(stderr, "\n");
    for (auto& m : v) cudaEventDestroy(m.ev);
    v.clear();
}

template <class T>
static int dalloc(gcpb200_ctx* c, T** p, size_t n, bool zero = true) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) {
        gcp_set_error("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
        return -1;
    }
    if (zero) cudaMemset(q, 0, n * sizeof(T));
    c->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return 0;
}
static int make_buf(gcpb200_ctx* c, DevBuf* b, size_t rows, int ld) {
    b->rows = rows;
    b->ld = ld;
    CHECK(dalloc(c, &b->p, rows * ld));
    return make_tmap_bf16(&b->map, b->p, rows, ld, ld, 128);
}

// ---------------------------------------------------------------------------------------------
// weight access + packing (host)
// ---------------------------------------------------------------------------------------------
struct WStore {
    std::map<std::string, const gcpb200_tensor*> m;
    const gcpb200_tensor* get(const std::string& k, int ndim) const {
        auto it = m.find(k);
        if (it == m.end()) {
            gcp_set_error("missing weight tensor '%s'", k.c_str());
            return nullptr;
        }
        if (it->second->ndim != ndim) {
            gcp_set_error("weight tensor '%s' has ndim %d, expected %d", k.c_str(), it->second->ndim, ndim);
            return nullptr;
        }
        return it->second;
    }
};

static int upload_mat(gcpb200_ctx* c, DevMat* d, int N, int K, const std::function<float(int, int)>& f,
                      const std::function<float(int)>& bias) {
    std::vector<bf16> h((size_t)N * K);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) h[(size_t)n * K + k] = __float2bfloat16(f(n, k));
    d->N = N;
    d->K = K;
    CHECK(dalloc(c, &d->w, (size_t)N * K, false));
    GCP_CUDA_CHECK(cudaMemcpy(d->w, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    std::vector<float> hb(N, 0.f);
    if (bias)
        for (int n = 0; n < N; ++n) hb[n] = bias(n);
    CHECK(dalloc(c, &d->bias, N, false));
    GCP_CUDA_CHECK(cudaMemcpy(d->bias, hb.data(), N * 4, cudaMemcpyHostToDevice));
    for (int i = 0; i < 5; ++i)
        if (N % (16 << i) == 0) CHECK(make_tmap_bf16(&d->map_box[i], d->w, N, K, K, 16 << i));
    return 0;
}
static int upload_f32(gcpb200_ctx* c, float** d, const std::vector<float>& h) {
    CHECK(dalloc(c, d, h.size(), false));
    GCP_CUDA_CHECK(cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}
static int upload_f32(gcpb200_ctx* c, const float** d, const std::vector<float>& h) {
    float* p;
    CHECK(upload_f32(c, &p, h));
    *d = p;
    return 0;
}

// linear weight element (o,i) of a Predictor layer: centre tap for the conv builder
struct Lin {
    const float* w = nullptr;
    const float* b = nullptr;
    int O = 0, I = 0, stride = 1, off = 0;
    float at(int o, int i) const { return (o < O && i < I) ? w[((size_t)o * I + i) * stride + off] : 0.f; }
    float bias(int o) const { return (b != nullptr && o < O) ? b[o] : 0.f; }
};
static int get_lin(const WStore& ws, const std::string& prefix, bool conv, bool has_bias, Lin* l) {
    const gcpb200_tensor* t = ws.get(prefix + (conv ? ".conv.weight" : ".linear.weight"), conv ? 4 : 2);
    if (!t) return -1;
    l->w = t->data;
    l->O = (int)t->shape[0];
    l->I = (int)t->shape[1];
    l->stride = conv ? 9 : 1;   // [O][I][3][3]: centre tap (1,1) is the only one that touches a 1x1 map
    l->off = conv ? 4 : 0;
    l->b = nullptr;
    if (has_bias) {
        const gcpb200_tensor* b = ws.get(prefix + (conv ? ".conv.bias" : ".linear.bias"), 1);
        if (!b) return -1;
        l->b = b->data;
    }
    return 0;
}

// head_perm: packed row -> original row (or -1 for a zero row); null = identity
static int pack_mlp(gcpb200_ctx* c, const WStore& ws, const std::string& prefix, bool conv, int d_in, int mid,
                    int d_out, int head_N, const std::function<int(int)>& head_perm, Mlp* m, DevMat* head2 = nullptr,
                    int head2_row0 = 0, int n_mid = 3) {
    Lin in, md[3], hd;
    m->n_mid = n_mid;
    CHECK(get_lin(ws, prefix + ".input", conv, true, &in));
    if (in.I != d_in || in.O != mid) {
        gcp_set_error("%s.input: got [%d,%d], expected [%d,%d]", prefix.c_str(), in.O, in.I, mid, d_in);
        return -1;
    }
    m->mid_k = mid < 64 ? 64 : mid;
    m->mid_valid = m->mid_k;
    m->gn_group = mid / 8;
    m->n_out = d_out;
    CHECK(upload_mat(c, &m->in, 128, d_in, [&](int n, int k) { return in.at(n, k); }, [&](int n) { return in.bias(n); }));
    for (int i = 0; i < n_mid; ++i) {
        const std::string p = prefix + ".pyramid-" + std::to_string(i);
        CHECK(get_lin(ws, p, conv, false, &md[i]));
        CHECK(upload_mat(c, &m->mid[i], 128, m->mid_k, [&](int n, int k) { return md[i].at(n, k); }, nullptr));
        const gcpb200_tensor* g = ws.get(p + ".norm.weight", 1);
        const gcpb200_tensor* b = ws.get(p + ".norm.bias", 1);
        if (!g || !b) return -1;
        std::vector<float> hg(128, 0.f), hb(128, 0.f);
        for (int k = 0; k < mid; ++k) {
            hg[k] = g->data[k];
            hb[k] = b->data[k];
        }
        CHECK(upload_f32(c, &m->gam[i], hg));
        CHECK(upload_f32(c, &m->bet[i], hb));
    }
    CHECK(get_lin(ws, prefix + ".head", conv, true, &hd));
    if (hd.O != d_out) {
        gcp_set_error("%s.head: got %d outputs, expected %d", prefix.c_str(), hd.O, d_out);
        return -1;
    }
    auto orig = [&](int n) { return head_perm ? head_perm(n) : n; };
    CHECK(upload_mat(c, &m->head, head_N, m->mid_k,
                     [&](int n, int k) { int o = orig(n); return o < 0 ? 0.f : hd.at(o, k); },
                     [&](int n) { int o = orig(n); return o < 0 ? 0.f : hd.bias(o); }));
    if (head2 != nullptr)
        CHECK(upload_mat(c, head2, head_N, m->mid_k, [&](int n, int k) { return hd.at(head2_row0 + n, k); },
                         [&](int n) { return hd.bias(head2_row0 + n); }));
    return 0;
}

// 1-D bilinear x2 matrix U[2n][n], align_corners=False (torch.nn.Upsample)
static std::vector<double> up_matrix(int n) {
    std::vector<double> U((size_t)2 * n * n, 0.0);
    for (int o = 0; o < 2 * n; ++o) {
        int i0, i1;
        float w0, w1;
        up2_src(o, n, i0, i1, w0, w1);
        U[(size_t)o * n + i0] += w0;
        U[(size_t)o * n + i1] += w1;
    }
    return U;
}
// Composite of bilinear x2 -> ZeroPad2d(1,2,1,2) -> conv k4 for one (co, ci) filter: dense [2n*2n][n*n]
// matrix C[(oy,ox)][(iy,ix)] = sum_{ky,kx} w[ky][kx] U[oy+ky-1][iy] U[ox+kx-1][ix] (U = 0 outside [0,2n)).
static void composite_filter(const float* w /*[4][4]*/, int n, const std::vector<double>& U, std::vector<double>& C) {
    const int m = 2 * n;
    std::vector<double> M((size_t)m * n * 4);   // M[oy][iy][kx] = sum_ky w[ky][kx] U[oy+ky-1][iy]
    for (int oy = 0; oy < m; ++oy)
        for (int iy = 0; iy < n; ++iy)
            for (int kx = 0; kx < 4; ++kx) {
                double s = 0;
                for (int ky = 0; ky < 4; ++ky) {
                    const int u = oy + ky - 1;
                    if (u >= 0 && u < m) s += (double)w[ky * 4 + kx] * U[(size_t)u * n + iy];
                }
                M[((size_t)oy * n + iy) * 4 + kx] = s;
            }
    C.assign((size_t)m * m * n * n, 0.0);
    for (int oy = 0; oy < m; ++oy)
        for (int ox = 0; ox < m; ++ox)
            for (int iy = 0; iy < n; ++iy)
                for (int ix = 0; ix < n; ++ix) {
                    double s = 0;
                    for (int kx = 0; kx < 4; ++kx) {
                        const int u = ox + kx - 1;
                        if (u >= 0 && u < m) s += M[((size_t)oy * n + iy) * 4 + kx] * U[(size_t)u * n + ix];
                    }
                    C[(((size_t)oy * m + ox) * n + iy) * n + ix] = s;
                }
}

// first input row of the 4-row K window of decoder layer 3's output band b (output rows 2b, 2b+1)
static int dec3_window_row0(int b) { return b < 1 ? 0 : (b - 1 > 4 ? 4 : b - 1); }

struct BNFold {
    std::vector<float> scale, shift;
};
static int fold_bn(const WStore& ws, const std::string& p, int C, BNFold* f) {
    const gcpb200_tensor *w = ws.get(p + ".weight", 1), *b = ws.get(p + ".bias", 1), *rm = ws.get(p + ".running_mean", 1),
                         *rv = ws.get(p + ".running_var", 1);
    if (!w || !b || !rm || !rv) return -1;
    f->scale.resize(C);
    f->shift.resize(C);
    for (int i = 0; i < C; ++i) {
        const float s = w->data[i] / sqrtf(rv->data[i] + 1e-5f);
        f->scale[i] = s;
        f->shift[i] = b->data[i] - rm->data[i] * s;
    }
    return 0;
}

// Decoder layers 1-3 as dense matrices.  fold = true: eval-mode BatchNorm folded into rows / bias (rollout path);
// fold = false: the raw convolutions (training path: batch statistics are applied by bn_stats / bn_apply kernels).
static int pack_decoder_dense(gcpb200_ctx* c, const WStore& ws, bool fold, DevMat* d1, DevMat* d2x, DevMat* d2s, DevMat* d3) {
    const std::string p = "decoder.net.net.";
    auto get_bn = [&](const std::string& name, int C, BNFold* f) -> int {
        if (fold) return fold_bn(ws, name, C, f);
        f->scale.assign(C, 1.0f);
        f->shift.assign(C, 0.0f);
        return 0;
    };
    // ---- layer 1: ConvTranspose2d(128->64,k4) on a 1x1 map + BN + ReLU == Linear 128 -> 64*16
    const gcpb200_tensor* w1 = ws.get(p + "net.conv.weight", 4);
    if (!w1) return -1;
    BNFold bn1, bn2, bn3;
    CHECK(get_bn(p + "net.norm", 64, &bn1));
    CHECK(upload_mat(c, d1, 1024, 128,
                     [&](int n, int k) { return w1->data[((size_t)k * 64 + n / 16) * 16 + n % 16] * bn1.scale[n / 16]; },
                     [&](int n) { return bn1.shift[n / 16]; }));
    // ---- layer 2: cat(x1 64ch, skip 64ch) 4x4 -> up -> pad -> conv(128->32) -> BN -> ReLU, as dense maps
    const gcpb200_tensor* w2 = ws.get(p + "pyramid-1.conv.weight", 4);
    if (!w2) return -1;
    CHECK(get_bn(p + "pyramid-1.norm", 32, &bn2));
    {
        const std::vector<double> U = up_matrix(4);
        // output column order of layer 2 = K order of layer 3: n = iy*256 + co*8 + ix (row-major bands of the 8x8 map),
        // so that a band of layer-3 output rows reads a contiguous K window (see layer 3)
        std::vector<float> Wx((size_t)2048 * 1024), Wsk((size_t)2048 * 1024);
        std::vector<double> C;
        auto col2 = [](int co, int o) { return (o >> 3) * 256 + co * 8 + (o & 7); };
        for (int co = 0; co < 32; ++co)
            for (int ci = 0; ci < 128; ++ci) {
                composite_filter(w2->data + ((size_t)co * 128 + ci) * 16, 4, U, C);
                std::vector<float>& dst = ci < 64 ? Wx : Wsk;
                const int cil = ci & 63;
                for (int o = 0; o < 64; ++o)
                    for (int i = 0; i < 16; ++i)
                        dst[(size_t)col2(co, o) * 1024 + cil * 16 + i] = (float)(C[(size_t)o * 16 + i] * bn2.scale[co]);
            }
        CHECK(upload_mat(c, d2x, 2048, 1024, [&](int n, int k) { return Wx[(size_t)n * 1024 + k]; }, nullptr));
        CHECK(upload_mat(c, d2s, 2048, 1024, [&](int n, int k) { return Wsk[(size_t)n * 1024 + k]; },
                         [&](int n) { return bn2.shift[(n & 255) >> 3]; }));
    }
    // ---- layer 3: 32ch 8x8 -> up -> pad -> conv(32->16) -> BN -> ReLU, as a banded dense map.  Output rows 2b, 2b+1
    // of the 16x16 map only depend on input rows b-1 .. b+2, so each group of 256 output columns (plane, band b)
    // multiplies a K window of 4 input rows (1024 of the 2048 inputs): half the MACs of the full composite.
    // Output columns are in the plane layout the tail kernel reads: n = ((co>>3)*256 + oy*16 + ox)*8 + (co&7).
    const gcpb200_tensor* w3 = ws.get(p + "pyramid-0.conv.weight", 4);
    if (!w3) return -1;
    CHECK(get_bn(p + "pyramid-0.norm", 16, &bn3));
    {
        const std::vector<double> U = up_matrix(8);
        std::vector<float> W3((size_t)4096 * 1024, 0.f);
        std::vector<double> C;
        double outside = 0.0;
        for (int co = 0; co < 16; ++co)
            for (int ci = 0; ci < 32; ++ci) {
                composite_filter(w3->data + ((size_t)co * 32 + ci) * 16, 8, U, C);
                for (int o = 0; o < 256; ++o) {
                    const size_t n3 = ((size_t)(co >> 3) * 256 + o) * 8 + (co & 7);
                    const int w0 = dec3_window_row0((o >> 4) >> 1);
                    for (int i = 0; i < 64; ++i) {
                        const int iy = i >> 3, ix = i & 7;
                        const double v = C[(size_t)o * 64 + i] * bn3.scale[co];
                        if (iy >= w0 && iy < w0 + 4) W3[n3 * 1024 + (iy - w0) * 256 + ci * 8 + ix] = (float)v;
                        else outside += fabs(v);
                    }
                }
            }
        if (outside != 0.0) {
            gcp_set_error("decoder layer 3: composite filter has weight outside its 4-row band (%g)", outside);
            return -1;
        }
        CHECK(upload_mat(c, d3, 4096, 1024, [&](int n, int k) { return W3[(size_t)n * 1024 + k]; },
                         [&](int n) { return bn3.shift[((n >> 3) / 256) * 8 + (n & 7)]; }));
    }
    return 0;
}

static int pack_decoder(gcpb200_ctx* c, const WStore& ws) {
    const std::string p = "decoder.net.net.";
    CHECK(pack_decoder_dense(c, ws, true, &c->dec1, &c->dec2x, &c->dec2s, &c->dec3));
    // ---- layers 4, 5: packed for the implicit-GEMM kernel + plain copies for the verification kernel
    const gcpb200_tensor* w4 = ws.get(p + "additional_conv_layer.conv.weight", 4);
    const gcpb200_tensor* b4 = ws.get(p + "additional_conv_layer.conv.bias", 1);
    const gcpb200_tensor* w5 = ws.get("decoder.net.gen_head.conv.weight", 4);
    const gcpb200_tensor* b5 = ws.get("decoder.net.gen_head.conv.bias", 1);
    if (!w4 || !b4 || !w5 || !b5) return -1;
    // Head channel c of the final convolution: DLM head = the 30 gen_head rows (15 mixture means + 15 scales); pixel-copy
    // head (adaptive model) = 3 gen_head rows followed by the 3 mask_head rows.
    const bool pc = c->model == GCPB200_MODEL_TREE_ADAPTIVE;
    const gcpb200_tensor *wm = nullptr, *bm = nullptr;
    if (pc) {
        wm = ws.get("decoder.net.mask_head.conv.weight", 4);
        bm = ws.get("decoder.net.mask_head.conv.bias", 1);
        if (!wm || !bm) return -1;
    }
    if (w5->shape[0] != (pc ? 3 : 30)) {
        gcp_set_error("decoder.net.gen_head.conv.weight has %d output channels, expected %d", (int)w5->shape[0], pc ? 3 : 30);
        return -1;
    }
    const int n_head = pc ? 6 : 30;
    auto head_w = [&](int co, int ci, int tap) -> float {
        if (co >= n_head) return 0.f;
        if (pc && co >= 3) return wm->data[((size_t)(co - 3) * 16 + ci) * 16 + tap];
        return w5->data[((size_t)co * 16 + ci) * 16 + tap];
    };
    auto head_b = [&](int co) -> float {
        if (co >= n_head) return 0.f;
        if (pc && co >= 3) return bm->data[co - 3];
        return b5->data[co];
    };
    {
        std::vector<bf16> h4(DT_W4_BYTES / 2), h5(DT_W5_BYTES / 2), p4(16 * 32 * 16), p5(32 * 16 * 16);
        for (int tap = 0; tap < 16; ++tap)
            for (int ci = 0; ci < 32; ++ci)
                for (int co = 0; co < 16; ++co) {
                    const float v = w4->data[((size_t)co * 32 + ci) * 16 + tap];
                    const int ks = ci >> 4, kc = (ci >> 3) & 1;
                    h4[(size_t)(tap * 2 + ks) * 256 + kc * 128 + co * 8 + (ci & 7)] = __float2bfloat16(v);
                    p4[((size_t)co * 32 + ci) * 16 + tap] = __float2bfloat16(v);
                }
        for (int tap = 0; tap < 16; ++tap)
            for (int ci = 0; ci < 16; ++ci)
                for (int co = 0; co < 32; ++co) {
                    const float v = head_w(co, ci, tap);
                    h5[(size_t)tap * 512 + (ci >> 3) * 256 + co * 8 + (ci & 7)] = __float2bfloat16(v);
                    p5[((size_t)co * 16 + ci) * 16 + tap] = __float2bfloat16(v);
                }
        // Toeplitz weight arrays of the quad-row kernel (dec_tail3.cuh): Z[ky][half h][row = 16 b + co][8 ci], block b
        // holds filter column kx = b - 3 (blocks 0-2 and 7-9 are zero).  conv 4: x-half of the input channels; conv 5:
        // the 15 mixture-mean channels, pre-scaled by 1/2 (sigmoid(v) = 0.5 + 0.5 tanh(v/2), exact in bf16).
        std::vector<bf16> z4(D3_W_BYTES / 2, __float2bfloat16(0.f)), z5(D3_W_BYTES / 2, __float2bfloat16(0.f));
        for (int ky = 0; ky < 4; ++ky)
            for (int h = 0; h < 2; ++h)
                for (int b = 3; b < 7; ++b)
                    for (int co = 0; co < 16; ++co)
                        for (int e = 0; e < 8; ++e) {
                            const size_t o = (size_t)ky * (D3_Z_KY / 2) + h * (D3_Z_CHUNK / 2) + (16 * b + co) * 8 + e;
                            const int ci = 8 * h + e, tap = ky * 4 + (b - 3);
                            z4[o] = __float2bfloat16(w4->data[((size_t)co * 32 + ci) * 16 + tap]);
                            if (pc) { if (co < 6) z5[o] = __float2bfloat16(head_w(co, ci, tap)); }
                            else if (co < 15) z5[o] = __float2bfloat16(0.5f * head_w(co, ci, tap));
                        }
        CHECK(dalloc(c, &c->z4, z4.size(), false));
        CHECK(dalloc(c, &c->z5, z5.size(), false));
        GCP_CUDA_CHECK(cudaMemcpy(c->z4, z4.data(), z4.size() * 2, cudaMemcpyHostToDevice));
        GCP_CUDA_CHECK(cudaMemcpy(c->z5, z5.data(), z5.size() * 2, cudaMemcpyHostToDevice));
        std::vector<float> hb5h(16, 0.f);
        for (int i = 0; i < (pc ? 6 : 15); ++i) hb5h[i] = (pc ? 1.0f : 0.5f) * head_b(i);
        CHECK(upload_f32(c, &c->b5h, hb5h));
        if (!pc && n_head >= 30) {
            // raw-head variants for the training-phase NLL (dec_tail3_raw_kernel): channels 0-14 and 15-29, unscaled
            std::vector<bf16> zm(D3_W_BYTES / 2, __float2bfloat16(0.f)), zs(D3_W_BYTES / 2, __float2bfloat16(0.f));
            for (int ky = 0; ky < 4; ++ky)
                for (int h = 0; h < 2; ++h)
                    for (int b = 3; b < 7; ++b)
                        for (int co = 0; co < 15; ++co)
                            for (int e = 0; e < 8; ++e) {
                                const size_t o = (size_t)ky * (D3_Z_KY / 2) + h * (D3_Z_CHUNK / 2) + (16 * b + co) * 8 + e;
                                const int ci = 8 * h + e, tap = ky * 4 + (b - 3);
                                zm[o] = __float2bfloat16(head_w(co, ci, tap));
                                zs[o] = __float2bfloat16(head_w(15 + co, ci, tap));
                            }
            CHECK(dalloc(c, &c->z5m, zm.size(), false));
            CHECK(dalloc(c, &c->z5s, zs.size(), false));
            GCP_CUDA_CHECK(cudaMemcpy(c->z5m, zm.data(), zm.size() * 2, cudaMemcpyHostToDevice));
            GCP_CUDA_CHECK(cudaMemcpy(c->z5s, zs.data(), zs.size() * 2, cudaMemcpyHostToDevice));
            std::vector<float> hbm(16, 0.f), hbs(16, 0.f);
            for (int i = 0; i < 15; ++i) {
                hbm[i] = head_b(i);
                hbs[i] = head_b(15 + i);
            }
            CHECK(upload_f32(c, &c->b5m, hbm));
            CHECK(upload_f32(c, &c->b5s, hbs));
        }
        CHECK(dalloc(c, &c->w4, h4.size(), false));
        CHECK(dalloc(c, &c->w5, h5.size(), false));
        CHECK(dalloc(c, &c->w4p, p4.size(), false));
        CHECK(dalloc(c, &c->w5p, p5.size(), false));
        GCP_CUDA_CHECK(cudaMemcpy(c->w4, h4.data(), h4.size() * 2, cudaMemcpyHostToDevice));
        GCP_CUDA_CHECK(cudaMemcpy(c->w5, h5.data(), h5.size() * 2, cudaMemcpyHostToDevice));
        GCP_CUDA_CHECK(cudaMemcpy(c->w4p, p4.data(), p4.size() * 2, cudaMemcpyHostToDevice));
        GCP_CUDA_CHECK(cudaMemcpy(c->w5p, p5.data(), p5.size() * 2, cudaMemcpyHostToDevice));
        std::vector<float> hb4(b4->data, b4->data + 16), hb5(32, 0.f);
        for (int i = 0; i < n_head; ++i) hb5[i] = head_b(i);
        CHECK(upload_f32(c, &c->b4, hb4));
        CHECK(upload_f32(c, &c->b5, hb5));
    }
    return 0;
}

static int pack_encoder(gcpb200_ctx* c, const WStore& ws) {
    const std::string p = "encoder.net.net.";
    auto up = [&](const std::string& k, int ndim, const float** d) -> int {
        const gcpb200_tensor* t = ws.get(k, ndim);
        if (!t) return -1;
        size_t n = 1;
        for (int i = 0; i < ndim; ++i) n *= t->shape[i];
        return upload_f32(c, d, std::vector<float>(t->data, t->data + n));
    };
    CHECK(up(p + "input.conv.weight", 4, &c->enc.w0));
    CHECK(up(p + "input.conv.bias", 1, &c->enc.b0));
    CHECK(up(p + "pyramid-0.conv.weight", 4, &c->enc.w1));
    CHECK(up(p + "pyramid-1.conv.weight", 4, &c->enc.w2));
    CHECK(up(p + "head.weight", 4, &c->enc.w3));
    CHECK(up(p + "head.bias", 1, &c->enc.b3));
    {   // k-major (k = ci*16 + ky*4 + kx) copies of the three strided convolutions for enc_train_*_kernel
        auto upT = [&](const std::string& k, int CO, int K, const float** d) -> int {
            const gcpb200_tensor* t = ws.get(k, 4);
            if (!t) return -1;
            std::vector<float> h((size_t)CO * K);
            for (int co = 0; co < CO; ++co)
                for (int kk = 0; kk < K; ++kk) h[(size_t)kk * CO + co] = t->data[(size_t)co * K + kk];
            return upload_f32(c, d, h);
        };
        CHECK(upT(p + "pyramid-0.conv.weight", 32, 256, &c->enc_w1t));
        CHECK(upT(p + "pyramid-1.conv.weight", 64, 512, &c->enc_w2t));
        CHECK(upT(p + "head.weight", 128, 1024, &c->enc_w3t));
    }
    BNFold f1, f2;
    CHECK(fold_bn(ws, p + "pyramid-0.norm", 32, &f1));
    CHECK(fold_bn(ws, p + "pyramid-1.norm", 64, &f2));
    CHECK(upload_f32(c, &c->enc.sc1, f1.scale));
    CHECK(upload_f32(c, &c->enc.sh1, f1.shift));
    CHECK(upload_f32(c, &c->enc.sc2, f2.scale));
    CHECK(upload_f32(c, &c->enc.sh2, f2.shift));
    return 0;
}

// LSTMCell weights packed for the EPI_LSTM epilogue: [W_ih | W_hh] along K, rows gate-interleaved so that a
// 32-column accumulator chunk holds i,f,g,o of 8 hidden units: packed row p = tile*256 + chunk*32 + gate*8 + u
// <-> gate*H + tile*64 + chunk*8 + u.
static int pack_lstm_cell(gcpb200_ctx* c, const WStore& ws, const std::string& lp, int H, DevMat* d) {
    const gcpb200_tensor *wih = ws.get(lp + "weight_ih", 2), *whh = ws.get(lp + "weight_hh", 2),
                         *bih = ws.get(lp + "bias_ih", 1), *bhh = ws.get(lp + "bias_hh", 1);
    if (!wih || !whh || !bih || !bhh) return -1;
    if (wih->shape[0] != 4 * H || wih->shape[1] != H || whh->shape[1] != H) {
        gcp_set_error("%sweight_ih: expected [%d,%d]", lp.c_str(), 4 * H, H);
        return -1;
    }
    auto orig = [H](int p) {
        const int tile = p >> 8, chunk = (p >> 5) & 7, gate = (p >> 3) & 3, u = p & 7;
        return gate * H + tile * 64 + chunk * 8 + u;
    };
    return upload_mat(c, d, 4 * H, 2 * H,
                      [&](int n, int k) {
                          const int o = orig(n);
                          return k < H ? wih->data[(size_t)o * H + k] : whh->data[(size_t)o * H + k - H];
                      },
                      [&](int n) { const int o = orig(n); return bih->data[o] + bhh->data[o]; });
}

static int pack_level(gcpb200_ctx* c, const WStore& ws, int l) {
    LevelW& L = c->lvl[l];
    // untied_layers (UntiedLayersTree, untied_layers_tree.py:9-17): one TreeModule per level; otherwise a single one
    const std::string tm = c->tied ? std::string("tree_module.") : "tree_module.tree_modules." + std::to_string(l) + ".";
    // prior head: packed row p = tile*256 + chunk*32 + part*16 + u  <->  part*256 + tile*128 + chunk*16 + u
    auto reparam_perm = [](int p) {
        const int tile = p >> 8, chunk = (p >> 5) & 7, part = (p >> 4) & 1, u = p & 15;
        return part * 256 + tile * 128 + chunk * 16 + u;
    };
    CHECK(pack_mlp(c, ws, tm + "prior", true, 2 * NZ_ENC, NZ_MID, 2 * NZ_VAE, 512, reparam_perm, &L.prior));
    if (c->has_train)
        CHECK(pack_mlp(c, ws, tm + "inference.q", true, 3 * NZ_ENC, NZ_MID, 2 * NZ_VAE, 512, reparam_perm, &L.q));
    if (l == 0)
        CHECK(pack_mlp(c, ws, tm + "lstm_initializer.net", true, 2 * NZ_ENC + NZ_VAE, INIT_MID, 2 * STATE, STATE, nullptr,
                       &L.init, &L.init_head_r, STATE));
    const std::string sp = tm + "subgoal_pred.";
    // six parent-state projections, stacked [P0,P2,P4 | P1,P3,P5] so that h-parts come first
    {
        const gcpb200_tensor* P[6];
        const gcpb200_tensor* Pb[6];
        for (int k = 0; k < 6; ++k) {
            P[k] = ws.get(sp + "projections." + std::to_string(k) + ".weight", 2);
            Pb[k] = ws.get(sp + "projections." + std::to_string(k) + ".bias", 1);
            if (!P[k] || !Pb[k]) return -1;
        }
        auto src = [](int g) { return g < 3 ? 2 * g : 2 * (g - 3) + 1; };
        CHECK(upload_mat(c, &L.proj, 6 * HID, 2 * HID,
                         [&](int n, int k) { return P[src(n / HID)]->data[(size_t)(n % HID) * 2 * HID + k]; },
                         [&](int n) { return Pb[src(n / HID)]->data[n % HID]; }));
    }
    const gcpb200_tensor* we = ws.get(sp + "embed.weight", 2);
    const gcpb200_tensor* be = ws.get(sp + "embed.bias", 1);
    if (!we || !be) return -1;
    const int EIN = 4 * NZ_ENC + NZ_VAE;   // 768 = [e_l, e_r, z, e_0, e_g]
    // all 768 input columns in the reference's order [e_l, e_r, z, e_0, e_g]: the context columns are two more K segments
    // of the embed GEMM (ROW_CAND rows of the latent array).  Hoisting them into a per-candidate row bias saved a third of
    // this GEMM's MMAs and cost more than that in its epilogue: a row-bias read is 32 scattered 16-byte loads per warp
    // instruction, and the embed GEMM (K = 512) was epilogue-bound at 3x its memory time (profiles/r2aq_embed.txt)
    CHECK(upload_mat(c, &L.embed_main, HID, EIN, [&](int n, int k) { return we->data[(size_t)n * EIN + k]; },
                     [&](int n) { return be->data[n]; }));
    for (int i = 0; i < N_LSTM; ++i) CHECK(pack_lstm_cell(c, ws, sp + "lstm." + std::to_string(i) + ".", HID, &L.lstm[i]));
    const gcpb200_tensor* wo = ws.get(sp + "output.weight", 2);
    const gcpb200_tensor* bo = ws.get(sp + "output.bias", 1);
    if (!wo || !bo) return -1;
    CHECK(upload_mat(c, &L.out, NZ_ENC, HID, [&](int n, int k) { return wo->data[(size_t)n * HID + k]; },
                     [&](int n) { return bo->data[n]; }));
    return 0;
}


// Training-only tensors: BatchNorm affine terms (batch statistics are computed on the fly), the conv-1d inference
// encoder (blox/torch/subnetworks.py:135-147), decoder layers 1-3 without BN folding.  The posterior MLPs are packed
// per level in pack_level.
static int pack_train(gcpb200_ctx* c, const WStore& ws) {
    const char* bn_names[5] = {"encoder.net.net.pyramid-0.norm", "encoder.net.net.pyramid-1.norm", "decoder.net.net.net.norm",
                               "decoder.net.net.pyramid-1.norm", "decoder.net.net.pyramid-0.norm"};
    const int bn_c[5] = {32, 64, 64, 32, 16};
    for (int i = 0; i < 5; ++i) {
        const gcpb200_tensor *g = ws.get(std::string(bn_names[i]) + ".weight", 1), *b = ws.get(std::string(bn_names[i]) + ".bias", 1);
        if (!g || !b) return -1;
        CHECK(upload_f32(c, &c->bn_g[i], std::vector<float>(g->data, g->data + bn_c[i])));
        CHECK(upload_f32(c, &c->bn_b[i], std::vector<float>(b->data, b->data + bn_c[i])));
    }
    const char* ie_names[3] = {"inf_encoder.net.input.conv", "inf_encoder.net.pyramid-0.conv", "inf_encoder.net.head.conv"};
    for (int i = 0; i < 3; ++i) {
        const gcpb200_tensor* w = ws.get(std::string(ie_names[i]) + ".weight", 3);
        if (!w) return -1;
        const int cin = (int)w->shape[1];
        if (w->shape[0] != 128 || w->shape[2] != 3 || cin != (i == 0 ? 129 : 128)) {
            gcp_set_error("%s.weight: unexpected shape", ie_names[i]);
            return -1;
        }
        std::vector<float> wt((size_t)3 * cin * 128);
        for (int k = 0; k < 3; ++k)
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < 128; ++co) wt[((size_t)k * cin + ci) * 128 + co] = w->data[((size_t)co * cin + ci) * 3 + k];
        CHECK(upload_f32(c, &c->ie_w[i], wt));
        if (i != 1) {
            const gcpb200_tensor* b = ws.get(std::string(ie_names[i]) + ".bias", 1);
            if (!b) return -1;
            CHECK(upload_f32(c, &c->ie_b[i], std::vector<float>(b->data, b->data + 128)));
        }
    }
    const gcpb200_tensor *gg = ws.get("inf_encoder.net.pyramid-0.norm.weight", 1), *gb = ws.get("inf_encoder.net.pyramid-0.norm.bias", 1);
    if (!gg || !gb) return -1;
    CHECK(upload_f32(c, &c->ie_gn_g, std::vector<float>(gg->data, gg->data + 128)));
    CHECK(upload_f32(c, &c->ie_gn_b, std::vector<float>(gb->data, gb->data + 128)));
    CHECK(pack_decoder_dense(c, ws, false, &c->dec1t, &c->dec2xt, &c->dec2st, &c->dec3t));
    return 0;
}

// VRNNCell weights of the sequential model (dense_rec.lstm.cell.{prior,gen_lstm}); inf_lstm / inf are not on the
// rollout path.
static int pack_sequential(gcpb200_ctx* c, const WStore& ws) {
    SeqW& S = c->seqw;
    const std::string cell = "dense_rec.lstm.cell.", g = cell + "gen_lstm.";
    const int H = c->lstm_hid;
    auto reparam_perm = [](int p) {
        const int tile = p >> 8, chunk = (p >> 5) & 7, part = (p >> 4) & 1, u = p & 15;
        return part * 256 + tile * 128 + chunk * 16 + u;
    };
    CHECK(pack_mlp(c, ws, cell + "prior", true, NZ_ENC, NZ_MID, 2 * NZ_VAE, 512, reparam_perm, &S.prior));
    // init_module head: reference columns [h0 c0 h1 c1 h2 c2] (var2state) -> packed [h0 h1 h2 | c0 c1 c2]
    auto init_perm = [H](int p) {
        const int is_c = p >= 3 * H, q = p - (is_c ? 3 * H : 0), layer = q / H, u = q % H;
        return (2 * layer + is_c) * H + u;
    };
    CHECK(pack_mlp(c, ws, g + "init_module", false, 2 * NZ_ENC, NZ_MID, 6 * H, 6 * H, init_perm, &S.init, nullptr, 0, 1));
    const gcpb200_tensor* we = ws.get(g + "embed.weight", 2);
    const gcpb200_tensor* be = ws.get(g + "embed.bias", 1);
    if (!we || !be) return -1;
    const int EIN = 3 * NZ_ENC + NZ_VAE;   // 640 = [x, z, e_0, e_g]
    if (we->shape[0] != H || we->shape[1] != EIN) {
        gcp_set_error("gen_lstm.embed.weight: expected [%d,%d]", H, EIN);
        return -1;
    }
    CHECK(upload_mat(c, &S.embed_main, H, NZ_ENC + NZ_VAE, [&](int n, int k) { return we->data[(size_t)n * EIN + k]; }, nullptr));
    CHECK(upload_mat(c, &S.embed_ctx, H, 2 * NZ_ENC,
                     [&](int n, int k) { return we->data[(size_t)n * EIN + NZ_ENC + NZ_VAE + k]; },
                     [&](int n) { return be->data[n]; }));
    for (int i = 0; i < N_LSTM; ++i) CHECK(pack_lstm_cell(c, ws, g + "lstm." + std::to_string(i) + ".", H, &S.lstm[i]));
    const gcpb200_tensor* wo = ws.get(g + "output.weight", 2);
    const gcpb200_tensor* bo = ws.get(g + "output.bias", 1);
    if (!wo || !bo) return -1;
    CHECK(upload_mat(c, &S.out, NZ_ENC, H, [&](int n, int k) { return wo->data[(size_t)n * H + k]; },
                     [&](int n) { return bo->data[n]; }));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM call builder
// ---------------------------------------------------------------------------------------------
struct Seg {
    const DevBuf* buf;
    int col0, k_len, mode, row_base;
    int group_cols;
    int group_col[16];
};
static Seg seg(const DevBuf& b, int col0, int k_len, int mode = ROW_LEVEL, int row_base = 0) {
    Seg s;
    memset(&s, 0, sizeof(s));
    s.buf = &b; s.col0 = col0; s.k_len = k_len; s.mode = mode; s.row_base = row_base;
    return s;
}
struct GemmDyn {          // device-side row count of a launch (GemmArgs::rows_dev / rows_dev_base)
    const int* rows_dev;
    int base;
};
static int gemm(gcpb200_ctx* c, cudaStream_t st, int rows, LevelGeom g, const std::vector<Seg>& segs, const DevMat& W,
                int BN, int epi, const EpiParams& ep, int w_row0 = 0, int n_cols = -1, const GemmDyn* dyn = nullptr) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.n_seg = (int)segs.size();
    int K = 0;
    for (int i = 0; i < a.n_seg; ++i) {
        const Seg& s = segs[i];
        a.a_map[i] = s.buf->map;
        a.seg[i].ptr = s.buf->p;
        a.seg[i].ld = s.buf->ld;
        a.seg[i].col0 = s.col0;
        a.seg[i].k_len = s.k_len;
        a.seg[i].row_mode = s.mode;
        a.seg[i].row_base = s.row_base;
        a.seg[i].group_cols = s.group_cols;
        for (int q = 0; q < 16; ++q) a.seg[i].group_col[q] = s.group_col[q];
        K += s.k_len;
    }
    if (K != W.K) {
        gcp_set_error("gemm: K mismatch (segments %d, weights %d)", K, W.K);
        return -1;
    }
    (void)w_row0;
    const int cluster = c->use_ref ? 1 : gemm_cluster_size(rows, c->max_cluster);
    const int box = BN / cluster;
    int bi = 0;
    while ((16 << bi) < box) ++bi;
    a.w_map = W.map_box[bi];
    a.w = W.w;
    a.w_ld = W.K;
    a.rows = rows;
    a.N = n_cols < 0 ? W.N : n_cols;
    a.K = K;
    a.g = g;
    if (dyn != nullptr) {
        a.rows_dev = dyn->rows_dev;
        a.rows_dev_base = dyn->base;
    }
    a.epi = ep;
    a.epi.bias = W.bias;
    ++c->launches;
    return launch_gemm(a, BN, epi, c->use_ref, st, c->sms, cluster);
}

static EpiParams epi_linear(int act, bf16* ob, int ob_ld, float* of, int of_ld, int n_valid, int ob_mode = ROW_LEVEL,
                            int of_mode = ROW_LEVEL) {
    EpiParams e;
    memset(&e, 0, sizeof(e));
    e.act = act;
    e.out_bf16 = ob; e.out_bf16_ld = ob_ld; e.out_bf16_mode = ob_mode;
    e.out_f32 = of; e.out_f32_ld = of_ld; e.out_f32_mode = of_mode;
    e.n_valid = n_valid;
    return e;
}

// The whole body (in + 3 GroupNorm layers) as one mlp_fused_kernel launch; result in c->tb like the unfused path.
// GCPB200_NO_FUSED_MLP=1 keeps the four-launch path (A/B measurements).
static bool fused_mlp_enabled() {
#if GCP_VERIFY
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("GCPB200_NO_FUSED_MLP");
        on = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return on == 1;
#else
    return true;
#endif
}
static int mlp_body_fused(gcpb200_ctx* c, cudaStream_t st, const Mlp& m, int rows, LevelGeom g, const std::vector<Seg>& in,
                          const GemmDyn* dyn = nullptr) {
    static bool configured_dev[64] = {false};     // function attributes are per device
    int dev = 0;
    GCP_CUDA_CHECK(cudaGetDevice(&dev));
    dev &= 63;
    if (!configured_dev[dev]) {
        GCP_CUDA_CHECK(cudaFuncSetAttribute(mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MLPF_SMEM_BYTES));
        configured_dev[dev] = true;
    }
    MlpFusedArgs a;
    memset(&a, 0, sizeof(a));
    GemmArgs& ia = a.in;
    ia.n_seg = (int)in.size();
    int K = 0;
    for (int i = 0; i < ia.n_seg; ++i) {
        const Seg& s = in[i];
        ia.a_map[i] = s.buf->map;
        ia.seg[i].ptr = s.buf->p;
        ia.seg[i].ld = s.buf->ld;
        ia.seg[i].col0 = s.col0;
        ia.seg[i].k_len = s.k_len;
        ia.seg[i].row_mode = s.mode;
        ia.seg[i].row_base = s.row_base;
        K += s.k_len;
    }
    if (K != m.in.K || rows % GEMM_BM || K % GEMM_BK) {
        gcp_set_error("fused mlp: bad shape rows %d K %d (weights K %d)", rows, K, m.in.K);
        return -1;
    }
    ia.w_map = m.in.map_box[3];
    ia.rows = rows;
    ia.N = 128;
    ia.K = K;
    ia.g = g;
    if (dyn != nullptr) {
        ia.rows_dev = dyn->rows_dev;
        ia.rows_dev_base = dyn->base;
    }
    for (int l = 0; l < 3; ++l) {
        a.w_mid[l] = m.mid[l].map_box[3];
        a.gam[l] = m.gam[l];
        a.bet[l] = m.bet[l];
    }
    a.bias_in = m.in.bias;
    a.out = c->tb.p;
    a.out_ld = c->tb.ld;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int tiles = rows / GEMM_BM;
    cfg.gridDim = dim3(tiles < c->sms ? tiles : c->sms);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = MLPF_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gemm_pdl_enabled() ? 1 : 0;
    ++c->launches;
    GCP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_fused_kernel, a));
    return 0;
}

// runs in -> mid x3; the last activation ends in c->tb (K = m.mid_k columns valid)
static int mlp_body(gcpb200_ctx* c, cudaStream_t st, const Mlp& m, int rows, LevelGeom g, const std::vector<Seg>& in,
                    const GemmDyn* dyn = nullptr) {
    bool plain = in.size() <= GEMM_MAX_SEGS;
    for (const Seg& s : in) plain = plain && s.group_cols == 0 && s.k_len % GEMM_BK == 0;
    if (!c->use_ref && fused_mlp_enabled() && plain && m.n_mid == 3 && m.mid_valid == 128 && m.mid_k == 128 &&
        m.gn_group == 16 && m.in.N == 128 && c->tb.ld == 128)
        return mlp_body_fused(c, st, m, rows, g, in, dyn);
    CHECK(gemm(c, st, rows, g, in, m.in, 128, EPI_LINEAR, epi_linear(ACT_LRELU, c->ta.p, c->ta.ld, nullptr, 0, m.mid_valid), 0, -1, dyn));
    DevBuf* src = &c->ta;
    DevBuf* dst = &c->tb;
    LevelGeom flat = g;
    for (int i = 0; i < m.n_mid; ++i) {
        EpiParams e = epi_linear(ACT_LRELU, dst->p, dst->ld, nullptr, 0, m.mid_valid);
        e.gn_gamma = m.gam[i];
        e.gn_beta = m.bet[i];
        e.gn_group = m.gn_group;
        CHECK(gemm(c, st, rows, flat, {seg(*src, 0, m.mid_k)}, m.mid[i], 128, EPI_GN, e, 0, -1, dyn));
        std::swap(src, dst);
    }
    // after an odd number of swaps the result is in `src` == tb
    return 0;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int gcpb200_create(gcpb200_ctx** out, const gcpb200_config* cfg) {
    if (!out || !cfg || cfg->max_candidates <= 0) {
        gcp_set_error("gcpb200_create: bad arguments");
        return -1;
    }
    GCP_CUDA_CHECK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    GCP_CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        gcp_set_error("gcpb200 needs a Blackwell sm_100 device, found %s (cc %d.%d)", prop.name, prop.major, prop.minor);
        return -1;
    }
    gcpb200_ctx* c = new gcpb200_ctx();
    c->cfg = *cfg;
    c->sms = prop.multiProcessorCount;
#if GCP_VERIFY
    c->use_ref = cfg->reserved0 != 0;
#else
    if (cfg->reserved0 != 0) {
        gcp_set_error("gcpb200_create: reserved0 must be 0");
        delete c;
        return -1;
    }
#endif
    c->Bp_max = (cfg->max_candidates + 127) / 128 * 128;
    c->slot_chunk = cfg->decoder_slot_chunk > 0 ? cfg->decoder_slot_chunk : 64;
    c->model = cfg->model;
    if (c->model != GCPB200_MODEL_TREE && c->model != GCPB200_MODEL_SEQUENTIAL && c->model != GCPB200_MODEL_TREE_ADAPTIVE) {
        gcp_set_error("gcpb200_create: unknown model kind %d", cfg->model);
        delete c;
        return -1;
    }
    const bool is_seq = c->model == GCPB200_MODEL_SEQUENTIAL;
    c->depth = cfg->hierarchy_levels > 0 ? cfg->hierarchy_levels : DEPTH;
    c->max_len = cfg->max_seq_len > 0 ? cfg->max_seq_len : MAX_LEN;
    c->tied = cfg->tied_layers != 0;
    c->n_nodes = (1 << c->depth) - 1;
    if (c->depth < 2 || c->depth > DEPTH || c->max_len < 3 || c->max_len > 256 || c->max_len % 4 != 0 ||
        (!is_seq && c->max_len > c->n_nodes)) {
        gcp_set_error("gcpb200_create: unsupported tree shape (hierarchy_levels %d in [2,8], max_seq_len %d a multiple of 4, <= 256 "
                      "and <= 2^levels - 1)", c->depth, c->max_len);
        delete c;
        return -1;
    }
    // latent rows are [slot][candidate]: tree = start, the in-order nodes, goal; sequential = frames 0..max_len-1, goal
    c->n_slots = is_seq ? c->max_len + 1 : c->n_nodes + 2;
    c->lstm_hid = is_seq ? 2 * HID : HID;
#if GCP_VERIFY
    if (const char* e = getenv("GCPB200_GEMM_CLUSTER")) c->max_cluster = atoi(e) > 0 ? atoi(e) : 1;
    if (const char* e = getenv("GCPB200_SEQ_GRAPH")) c->seq_graph_on = atoi(e) != 0;
#endif
    const size_t Bp = c->Bp_max, NL = is_seq ? Bp : ((size_t)1 << (c->depth - 1)) * Bp, NS = (size_t)c->n_slots * Bp, ND = 256 * Bp;
    int rc = 0;
    rc |= dalloc(c, &c->lat_f32, NS * NZ_ENC);
    rc |= make_buf(c, &c->lat, NS, NZ_ENC);
    if (is_seq) {
        rc |= make_buf(c, &c->hs[0], Bp, 3 * c->lstm_hid);
        rc |= make_buf(c, &c->hs[1], Bp, 3 * c->lstm_hid);
        rc |= dalloc(c, &c->cs, Bp * 3 * c->lstm_hid);
        rc |= make_buf(c, &c->xa, NL, c->lstm_hid);
    } else {
        rc |= make_buf(c, &c->hid, NS, STATE);
        rc |= make_buf(c, &c->xa, NL, HID);
        rc |= make_buf(c, &c->xb, NL, HID);
        rc |= make_buf(c, &c->sh, NL, 6 * HID);    // projected parent state: [h0 h1 h2 | c0 c1 c2]
    }
    rc |= make_buf(c, &c->zeta, NL, NZ_VAE);
    rc |= make_buf(c, &c->ta, ND, 128);
    rc |= make_buf(c, &c->tb, ND, 128);
    rc |= make_buf(c, &c->s2b, Bp, 1024);
    rc |= make_buf(c, &c->x1, (size_t)c->slot_chunk * Bp, 1024);
    rc |= make_buf(c, &c->x2, (size_t)c->slot_chunk * Bp, 2048);
    rc |= make_buf(c, &c->x3, (size_t)c->slot_chunk * Bp, 4096);
    c->pair_rows = (int)((c->model == GCPB200_MODEL_TREE_ADAPTIVE ? c->n_nodes : c->max_len + 1) * Bp + 256);
    rc |= make_buf(c, &c->pairs, (size_t)c->pair_rows, 256);
    rc |= make_buf(c, &c->seqb, (size_t)(c->max_len + 1) * Bp + 256, NZ_ENC);
    rc |= dalloc(c, &c->ctxb, Bp * c->lstm_hid);
    rc |= dalloc(c, &c->logits, Bp * 256);
    rc |= dalloc(c, &c->s0, Bp * 4096);
    rc |= dalloc(c, &c->s2, Bp * 1024);
    rc |= dalloc(c, &c->rowbias2, Bp * 2048);
    rc |= dalloc(c, &c->skip_up, Bp * 2 * DT_PSTRIDE * 8);
    rc |= dalloc(c, &c->s4, Bp * 256 * 64);
    rc |= dalloc(c, &c->exist_slot, ND);
    rc |= dalloc(c, &c->e_df, Bp * c->n_nodes * NZ_ENC);
    rc |= dalloc(c, &c->seq, Bp * c->max_len * NZ_ENC);
    rc |= dalloc(c, &c->rowcost, (size_t)c->pair_rows * 2);
    rc |= dalloc(c, &c->goal_tail, 256);
    rc |= dalloc(c, &c->end_ind, Bp);
    rc |= dalloc(c, &c->scratch_ei, 8);
    rc |= dalloc(c, &c->scratch_given, Bp);
    rc |= dalloc(c, &c->frame_node, Bp * 256);
    if (c->model == GCPB200_MODEL_TREE) {
        const size_t kept_cap = (size_t)c->max_len * Bp + 256;
        rc |= make_buf(c, &c->latc, kept_cap, NZ_ENC);
        rc |= dalloc(c, &c->row_cand, kept_cap);
        rc |= dalloc(c, &c->row_node, kept_cap);
        rc |= dalloc(c, &c->row_off, Bp + 1);
        rc |= dalloc(c, &c->n_rows, 1);
        rc |= dalloc(c, &c->frame_sq, 256 * Bp);
        rc |= dalloc(c, &c->tree_tiles, (size_t)c->n_nodes * (Bp >> 7) + 2 * DEPTH + 2);
        rc |= dalloc(c, &c->tree_rows, DEPTH);
    }
    rc |= dalloc(c, &c->refit_part, (size_t)REFIT_SPLITS * c->n_nodes * NZ_VAE * 2, false);
    if (rc) {
        gcpb200_destroy(c);
        return -1;
    }
    {
        cudaError_t ce = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_copy_start, cudaEventDisableTiming);
        for (int i = 0; i < 5 && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->prep_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_prep_fork, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_prep_join, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->proj_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_proj_fork, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_proj_join, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&c->seq_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_seq_fork, cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&c->ev_seq_join, cudaEventDisableTiming);
        if (ce != cudaSuccess) {
            gcp_set_error("copy stream / event creation failed: %s", cudaGetErrorString(ce));
            gcpb200_destroy(c);
            return -1;
        }
    }
    // An SM cannot hold CTAs of kernels that ask for different shared-memory carve-outs: without this the persistent
    // GEMM / decoder CTAs (200+ KB of shared memory) wait until every block of the upload kernel has left the SM and the
    // "overlapped" upload serialises with the rollout (measured: tree 6.2 -> 11.1 ms).  So the upload kernel asks for the
    // same (maximum) carve-out as those kernels, and it only uses 64 small blocks (see the launch-shape note in
    // gcpb200_rollout) so that the small kernels of the rollout that run with the default carve-out still find SMs they
    // can be scheduled on.
    cudaFuncSetAttribute(upload_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaError_t e = cudaFuncSetAttribute(dec_tail3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(enc_train_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENCA_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(enc_train_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENCB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(enc_train_c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENCC_SMEM);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dec_tail3_pc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM_BYTES);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dec_tail3_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D3_SMEM_BYTES);
#if GCP_VERIFY
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(dec_tail_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * DT_PLANE_BYTES + 128);
#endif
    if (e != cudaSuccess) {
        gcp_set_error("cudaFuncSetAttribute(dec_tail) failed: %s", cudaGetErrorString(e));
        gcpb200_destroy(c);
        return -1;
    }
    *out = c;
    return 0;
}

extern "C" void gcpb200_destroy(gcpb200_ctx* c) {
    if (!c) return;
    for (auto& g : c->seq_graphs) cudaGraphExecDestroy(g.exec);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
    if (c->ev_prep_fork) cudaEventDestroy(c->ev_prep_fork);
    if (c->ev_prep_join) cudaEventDestroy(c->ev_prep_join);
    if (c->proj_stream) cudaStreamDestroy(c->proj_stream);
    if (c->ev_proj_fork) cudaEventDestroy(c->ev_proj_fork);
    if (c->ev_proj_join) cudaEventDestroy(c->ev_proj_join);
    if (c->seq_stream) cudaStreamDestroy(c->seq_stream);
    if (c->ev_seq_fork) cudaEventDestroy(c->ev_seq_fork);
    if (c->ev_seq_join) cudaEventDestroy(c->ev_seq_join);
    if (c->ev_copy_start) cudaEventDestroy(c->ev_copy_start);
    for (int i = 0; i < 5; ++i)
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    for (void* p : c->allocs) cudaFree(p);
    if (c->topk_sel) cudaFree(c->topk_sel);
    delete c;
}

extern "C" int64_t gcpb200_launch_count(gcpb200_ctx* c) { return c ? c->launches : 0; }

extern "C" int gcpb200_profile_enable(gcpb200_ctx* c, int on) {
    if (!c) return -1;
    c->profile = on != 0;
    return 0;
}
// Synchronises on the recorded events, returns accumulated milliseconds per phase since the last call
// and the number of node images / launches the decoder-tail kernel processed; then resets.
extern "C" int gcpb200_profile_read(gcpb200_ctx* c, double* ms, int64_t* tail_images, int64_t* tail_launches) {
    if (!c || !ms) return -1;
    for (auto& sp : c->spans) {
        GCP_CUDA_CHECK(cudaEventSynchronize(sp.b));
        float t = 0.f;
        GCP_CUDA_CHECK(cudaEventElapsedTime(&t, sp.a, sp.b));
        c->prof_ms[sp.phase] += t;
        c->ev_pool.push_back(sp.a);
        c->ev_pool.push_back(sp.b);
    }
    c->spans.clear();
    for (int i = 0; i < GCPB200_N_PHASES; ++i) {
        ms[i] = c->prof_ms[i];
        c->prof_ms[i] = 0;
    }
    if (tail_images) *tail_images = c->prof_tail_images;
    if (tail_launches) *tail_launches = c->prof_tail_launches;
    c->prof_tail_images = c->prof_tail_launches = 0;
    return 0;
}

extern "C" int gcpb200_load_weights(gcpb200_ctx* c, const gcpb200_tensor* tensors, int n) {
    if (!c || !tensors) {
        gcp_set_error("gcpb200_load_weights: bad arguments");
        return -1;
    }
    GCP_CUDA_CHECK(cudaSetDevice(c->cfg.device));
    WStore ws;
    for (int i = 0; i < n; ++i) ws.m[tensors[i].name] = &tensors[i];
    CHECK(pack_encoder(c, ws));
    CHECK(pack_decoder(c, ws));
    // training-only tensors are optional: planner checkpoints stripped of them still load (forward_loss then refuses)
    c->has_train = c->model == GCPB200_MODEL_TREE && !c->tied && c->depth == DEPTH && c->max_len == MAX_LEN &&
                   ws.m.count("inf_encoder.net.input.conv.weight") != 0 &&
                   ws.m.count("tree_module.tree_modules.0.inference.q.input.conv.weight") != 0;
    if (c->has_train) CHECK(pack_train(c, ws));
    if (c->model == GCPB200_MODEL_SEQUENTIAL) {
        CHECK(pack_sequential(c, ws));
    } else {
        for (int l = 0; l < (c->tied ? 1 : c->depth); ++l) CHECK(pack_level(c, ws, l));
        if (c->model == GCPB200_MODEL_TREE_ADAPTIVE)
            CHECK(pack_mlp(c, ws, std::string(c->tied ? "tree_module." : "tree_module.tree_modules.0.") + "binding.distance_predictor", true, 2 * NZ_ENC, NZ_MID, 1, 128,
                           nullptr, &c->distance_pred));
        else
            CHECK(pack_mlp(c, ws, std::string(c->tied ? "tree_module." : "tree_module.tree_modules.0.") + "binding.existence_predictor", true, NZ_ENC, NZ_MID, 1, 128,
                           nullptr, &c->existence));
    }
    CHECK(pack_mlp(c, ws, "length_pred.p", true, 2 * NZ_ENC, NZ_MID, c->max_len, 256, nullptr, &c->length_pred));
    // auxiliary heads exist only when the model config attaches them (attach_inv_mdl / attach_state_regressor)
    c->has_inv = ws.m.count("inv_mdl.action_pred.input.linear.weight") != 0;
    c->has_state = ws.m.count("state_regressor.input.linear.weight") != 0;
    if (c->has_inv) CHECK(pack_mlp(c, ws, "inv_mdl.action_pred", false, 2 * NZ_ENC, 128, 2, 128, nullptr, &c->inv_mdl));
    if (c->has_state) CHECK(pack_mlp(c, ws, "state_regressor", false, NZ_ENC, NZ_MID, 2, 128, nullptr, &c->state_reg));
    c->has_cost = false;
    if (c->cfg.attach_cost_mdl) {
        CHECK(pack_mlp(c, ws, "cost_mdl.cost_pred", false, 2 * NZ_ENC, 128, 1, 128, nullptr, &c->cost_mdl));
        c->has_cost = true;
    }
    c->weights_loaded = true;
    return 0;
}

static int check_ready(gcpb200_ctx* c, int B) {
    if (!c) {
        gcp_set_error("null context");
        return -1;
    }
    if (!c->weights_loaded) {
        gcp_set_error("weights not loaded (call gcpb200_load_weights first)");
        return -1;
    }
    if (B <= 0 || B > c->Bp_max) {
        gcp_set_error("B = %d outside (0, max_candidates = %d]", B, c->Bp_max);
        return -1;
    }
    return 0;
}

#define LAUNCH_CHECK()                         \
    do {                                       \
        ++c->launches;                         \
        GCP_CUDA_CHECK(cudaGetLastError());    \
    } while (0)

static int compute_frame_map(gcpb200_ctx* c, const long long* end_ind, int B, cudaStream_t st) {
    GCP_CUDA_CHECK(cudaMemsetAsync(c->frame_node, 0, (size_t)B * c->max_len * sizeof(int), st));
    prune_map_kernel<<<(B * c->n_nodes + 255) / 256, 256, 0, st>>>(end_ind, B, c->depth, c->max_len, c->frame_node);
    LAUNCH_CHECK();
    return 0;
}

struct CommonIO {
    const float *I_0, *I_g;
    int images_shared;
    const int64_t* end_ind;
    uint64_t seed;
    int B;
    float *e_0, *e_g, *seq_len_logits;
    int64_t* end_ind_out;
    int sort_lengths;
};

// Encoder on the start / goal images (latent slots 0 and goal_row0 / Bp, decoder skips of I_0) and the rollout length
// (LengthPredictorModule + OneHotCategorical sample, or injected): BaseGCPModel.run_encoder + get_end_ind
// (gcp/prediction/models/base_gcp.py:184-229).  Shared by the tree and the sequential model.
static int run_encoder_length(gcpb200_ctx* c, cudaStream_t st, const CommonIO& in, int Bp, int goal_row0) {
    const CommonIO* io = &in;
    const int B = in.B;
    const LevelGeom flat = {Bp, 0, c->depth};
    // ---- 1. encoder on start / goal images -> latent slots 0 and 256 (+ decoder skips of I_0)
    const int n_img = io->images_shared ? 1 : B;
    if (io->images_shared) {
        // one start / goal pair: a cluster of 8 CTAs per image instead of one CTA (same work items, same bits)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * ENCC_CTAS);
        cfg.blockDim = dim3(ENCC_THREADS);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = ENCC_CTAS;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        GCP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, encoder_cluster_kernel, (const float*)io->I_0, (const float*)io->I_g, c->enc,
                                          c->lat_f32, c->lat.p, 0, goal_row0, c->s0, c->s2, c->s2b.p));
    } else {
        encoder_kernel<<<dim3(n_img, 2), ENC_THREADS, 0, st>>>(io->I_0, io->I_g, c->enc, c->lat_f32, c->lat.p, 0, goal_row0, c->s0,
                                                                c->s2, c->s2b.p);
    }
    LAUNCH_CHECK();
    if (io->images_shared) {
        broadcast_rows_kernel<<<(Bp * NZ_ENC + 255) / 256, 256, 0, st>>>(c->lat_f32, c->lat.p, 0, Bp, NZ_ENC);
        LAUNCH_CHECK();
        broadcast_rows_kernel<<<(Bp * NZ_ENC + 255) / 256, 256, 0, st>>>(c->lat_f32, c->lat.p, goal_row0, Bp, NZ_ENC);
        LAUNCH_CHECK();
    }
    if (io->e_0) GCP_CUDA_CHECK(cudaMemcpyAsync(io->e_0, c->lat_f32, (size_t)B * NZ_ENC * 4, cudaMemcpyDeviceToDevice, st));
    if (io->e_g)
        GCP_CUDA_CHECK(cudaMemcpyAsync(io->e_g, c->lat_f32 + (size_t)goal_row0 * NZ_ENC, (size_t)B * NZ_ENC * 4,
                                       cudaMemcpyDeviceToDevice, st));

    // ---- 2. rollout length (LengthPredictorModule + OneHotCategorical sample, or injected)
    const std::vector<Seg> ctx_in = {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, 0), seg(c->lat, 0, NZ_ENC, ROW_LEVEL, goal_row0)};
    if (io->seq_len_logits || !io->end_ind) {
        CHECK(mlp_body(c, st, c->length_pred, Bp, flat, ctx_in));
        CHECK(gemm(c, st, Bp, flat, {seg(c->tb, 0, c->length_pred.mid_k)}, c->length_pred.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->logits, 256, c->max_len)));
        if (io->seq_len_logits)
            GCP_CUDA_CHECK(cudaMemcpy2DAsync(io->seq_len_logits, c->max_len * 4, c->logits, 256 * 4, c->max_len * 4, B,
                                             cudaMemcpyDeviceToDevice, st));
    }
    if (io->end_ind) {
        GCP_CUDA_CHECK(cudaMemcpyAsync(c->end_ind, io->end_ind, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        sample_length_kernel<<<(B + 7) / 8, 256, 0, st>>>(c->logits, 256, c->max_len, B, io->seed, c->end_ind);
        LAUNCH_CHECK();
        if (io->sort_lengths && io->images_shared) {
            sort_lengths_desc_kernel<<<1, 1024, 0, st>>>(c->end_ind, B);
            LAUNCH_CHECK();
        }
    }
    if (io->end_ind_out) GCP_CUDA_CHECK(cudaMemcpyAsync(io->end_ind_out, c->end_ind, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));

    return 0;
}

// Decoder over latent slots 1 .. n_dec (DecoderModule.decode_seq, blox/torch/encoder_decoder.py:358-372): three dense
// composite layers as GEMMs + the implicit-GEMM tail kernel.  The image of (candidate c, slot s) goes to
// images[(c * n_layout + s - 1) * 3072].
// Per-call constants of the decoder: the up-sampled encoder skip, the skip half of the 32->16 tail conv in the quad
// layout of dec_tail3, and the skip half of the 128->32 conv as a per-candidate row bias (the conv is linear in its input).
static int decoder_prepare(gcpb200_ctx* c, cudaStream_t st, int images_shared, int B, int Bp) {
    const LevelGeom flat = {Bp, 0, c->depth};
    const int n_skip = images_shared ? 1 : B;
    skip_prep_kernel<<<n_skip, 256, 0, st>>>(c->s0, c->skip_up, n_skip);
    LAUNCH_CHECK();
    if (!c->use_ref) {
        skip_term3_kernel<<<dim3(8, n_skip), 128, 0, st>>>(c->skip_up, c->w4p, c->b4, c->s4);
        LAUNCH_CHECK();
    }
    CHECK(gemm(c, st, images_shared ? 128 : Bp, flat, {seg(c->s2b, 0, 1024)}, c->dec2s, 256, EPI_LINEAR,
               epi_linear(ACT_NONE, nullptr, 0, c->rowbias2, 2048, 2048)));
    return 0;
}

// Decodes the tree slots first, first + step, ..., `count` of them (step 1: a contiguous range; step 2 from slot 1: the
// nodes of tree level 7; step 4 from slot 2: level 6; step 4 from slot 4: levels 0-5), slot_chunk slots per pass: three dense GEMM layers on
// the slot-major latents, then the tail kernel writes image node = slot - 1 of every candidate.  step 2 needs the
// tcgen05 tail kernel (the SIMT verification kernel only takes contiguous ranges).
static int decoder_slots(gcpb200_ctx* c, cudaStream_t st, int images_shared, int B, int Bp, int first, int step, int count,
                         float* images, int n_layout, const float* I_0, const float* I_g, const float* l2_goal = nullptr) {
    const bool pc = c->model == GCPB200_MODEL_TREE_ADAPTIVE;   // pixel-copy head: needs the start / goal images
    const LevelGeom flat = {Bp, 0, c->depth};
    if (step != 1 && ((step != 2 && step != 4) || c->use_ref)) {
        gcp_set_error("decoder_slots: unsupported slot step %d", step);
        return -1;
    }
    for (int k0 = 0; k0 < count; k0 += c->slot_chunk) {
        ProfScope* dsc = new ProfScope(c, st, 2);
        const int ns = (k0 + c->slot_chunk <= count) ? c->slot_chunk : count - k0;
        const int s0 = first + k0 * step;
        const int rows = ns * Bp;
        if (step == 1) {
            CHECK(gemm(c, st, rows, flat, {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, s0 * Bp)}, c->dec1, 256, EPI_LINEAR,
                       epi_linear(ACT_RELU, c->x1.p, 1024, nullptr, 0, 1024)));
        } else {
            // row block j of the A operand = slot s0 + step * j: the geometry of tree level 7 (step 2) / 6 (step 4) maps
            // j to slot (2j + 1) * step / 2 (ROW_SELF); row_base shifts that to s0
            const int half = step / 2;
            const LevelGeom gl = {Bp, step == 2 ? c->depth - 1 : c->depth - 2, c->depth};
            CHECK(gemm(c, st, rows, gl, {seg(c->lat, 0, NZ_ENC, ROW_SELF, (s0 - half) * Bp)}, c->dec1, 256, EPI_LINEAR,
                       epi_linear(ACT_RELU, c->x1.p, 1024, nullptr, 0, 1024)));
        }
        {
            EpiParams e = epi_linear(ACT_RELU, c->x2.p, 2048, nullptr, 0, 2048);
            e.rowbias = c->rowbias2;
            e.rowbias_ld = images_shared ? 0 : 2048;
            CHECK(gemm(c, st, rows, flat, {seg(c->x1, 0, 1024)}, c->dec2x, 256, EPI_LINEAR, e));
        }
        {
            Seg a3 = seg(c->x2, 0, 1024);      // banded: column group (plane, band) reads its own 4-row K window
            a3.group_cols = 256;
            for (int q = 0; q < 16; ++q) a3.group_col[q] = dec3_window_row0(q & 7) * 256;
            CHECK(gemm(c, st, rows, flat, {a3}, c->dec3, 256, EPI_LINEAR,
                       epi_linear(ACT_RELU, c->x3.p, 4096, nullptr, 0, 4096)));
        }
        delete dsc;
        ProfScope tsc(c, st, 3);
        if (c->profile) {
            c->prof_tail_images += (long long)ns * B;
            ++c->prof_tail_launches;
        }
#if GCP_VERIFY
        if (c->use_ref) {
            DecTailArgs a;
            memset(&a, 0, sizeof(a));
            a.x3 = c->x3.p; a.skip_up = c->skip_up; a.skip_stride = images_shared ? 0 : 2 * DT_PSTRIDE * 8;
            a.w4 = c->w4; a.w5 = c->w5; a.b4 = c->b4; a.b5 = c->b5;
            a.images = images; a.Bp = Bp; a.n_cand = B; a.slot0 = s0; a.n_slots = ns; a.n_nodes = n_layout;
            a.head = pc; a.src0 = I_0; a.srcg = I_g; a.src_stride = images_shared ? 0 : 3072;
            dec_tail_ref_kernel<<<ns * B, 256, 6 * DT_PLANE_BYTES + 128, st>>>(a, c->w4p, c->w5p);
        } else
#endif
        {
            DecTail3Args a;
            memset(&a, 0, sizeof(a));
            a.x3 = c->x3.p; a.s4 = c->s4; a.s4_stride = images_shared ? 0 : 256 * 64;
            a.w4 = c->z4; a.w5 = c->z5; a.b5h = c->b5h;
            a.images = images; a.Bp = Bp; a.n_cand = B; a.slot0 = s0; a.n_slots = ns; a.n_nodes = n_layout;
            a.slot_extra = step - 1;
            if (l2_goal != nullptr) {       // fused L2 image cost: per-image sums of squares, [cand][node]
                a.frame_sq = c->frame_sq;
                a.l2_goal = l2_goal;
            }
            // one persistent CTA per SM; each takes a contiguous run of (candidate, slot) images
            const long long n_img = (long long)B * ns;
            a.src0 = I_0; a.srcg = I_g; a.src_stride = images_shared ? 0 : 3072;
            if (pc) dec_tail3_pc_kernel<<<(unsigned)(n_img < c->sms ? n_img : c->sms), D3_THREADS, D3_SMEM_BYTES, st>>>(a);
            else dec_tail3_kernel<<<(unsigned)(n_img < c->sms ? n_img : c->sms), D3_THREADS, D3_SMEM_BYTES, st>>>(a);
        }
        LAUNCH_CHECK();
    }
    return 0;
}

static int run_decoder(gcpb200_ctx* c, cudaStream_t st, int images_shared, int B, int Bp, int n_dec, float* images,
                       int n_layout, const float* I_0 = nullptr, const float* I_g = nullptr) {
    CHECK(decoder_prepare(c, st, images_shared, B, Bp));
    return decoder_slots(c, st, images_shared, B, Bp, 1, 1, n_dec, images, n_layout, I_0, I_g);
}

// Planner mode: decodes only the nodes balanced pruning keeps (DecoderModule.decode_seq restricted to the frames
// GCPImageSimulator.rollout hands to the cost, cem_simulator.py:29-61).  The kept (candidate, frame) pairs are compacted
// into consecutive latent rows (candidate-major); the number of rows is known to the device only, so the GEMMs and the tail
// kernel are launched for the capacity of a chunk and read the live row count from c->n_rows.  Needs c->frame_node
// (compute_frame_map), c->row_off / c->n_rows (kept_offsets_kernel), decoder_prepare and the finished tree.
static int decoder_kept(gcpb200_ctx* c, cudaStream_t st, int images_shared, int B, int Bp, float* images, const float* l2_goal) {
    const LevelGeom flat = {Bp, 0, c->depth};
    {
        const size_t n = (size_t)B * c->max_len * 16;
        gather_kept_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->lat.p, c->frame_node, c->end_ind, c->row_off, B, Bp,
                                                                           c->max_len, 1, c->latc.p, c->row_cand, c->row_node);
        LAUNCH_CHECK();
    }
    const int chunk_rows = c->slot_chunk * Bp;
    const int cap = c->max_len * B;                 // at most 200 kept frames per candidate
    for (int r0 = 0; r0 < cap; r0 += chunk_rows) {
        const int rows = std::min(chunk_rows, (cap - r0 + 255) / 256 * 256);
        {
            ProfScope dsc(c, st, 2);
            GemmDyn dyn = {c->n_rows, r0};
            CHECK(gemm(c, st, rows, flat, {seg(c->latc, 0, NZ_ENC, ROW_LEVEL, r0)}, c->dec1, 256, EPI_LINEAR,
                       epi_linear(ACT_RELU, c->x1.p, 1024, nullptr, 0, 1024), 0, -1, &dyn));
            {
                EpiParams e = epi_linear(ACT_RELU, c->x2.p, 2048, nullptr, 0, 2048);
                e.rowbias = c->rowbias2;
                e.rowbias_ld = images_shared ? 0 : 2048;
                e.rowbias_idx = images_shared ? nullptr : c->row_cand + r0;
                CHECK(gemm(c, st, rows, flat, {seg(c->x1, 0, 1024)}, c->dec2x, 256, EPI_LINEAR, e, 0, -1, &dyn));
            }
            Seg a3 = seg(c->x2, 0, 1024);
            a3.group_cols = 256;
            for (int q = 0; q < 16; ++q) a3.group_col[q] = dec3_window_row0(q & 7) * 256;
            CHECK(gemm(c, st, rows, flat, {a3}, c->dec3, 256, EPI_LINEAR, epi_linear(ACT_RELU, c->x3.p, 4096, nullptr, 0, 4096), 0, -1,
                       &dyn));
        }
        ProfScope tsc(c, st, 3);
        DecTail3Args a;
        memset(&a, 0, sizeof(a));
        a.x3 = c->x3.p; a.s4 = c->s4; a.s4_stride = images_shared ? 0 : 256 * 64;
        a.w4 = c->z4; a.w5 = c->z5; a.b5h = c->b5h;
        a.images = images; a.Bp = Bp; a.n_cand = B; a.n_slots = c->slot_chunk; a.n_nodes = c->n_nodes;
        a.n_img_dev = c->n_rows; a.row_cand = c->row_cand; a.row_node = c->row_node; a.img_base = r0;
        if (l2_goal != nullptr) {
            a.frame_sq = c->frame_sq;
            a.l2_goal = l2_goal;
        }
        const long long n_img = (long long)rows;
        dec_tail3_kernel<<<(unsigned)(n_img < c->sms ? n_img : c->sms), D3_THREADS, D3_SMEM_BYTES, st>>>(a);
        LAUNCH_CHECK();
        if (c->profile) ++c->prof_tail_launches;
    }
    if (c->profile) {       // measurement aid only: the live row count is on the device
        int n = 0;
        GCP_CUDA_CHECK(cudaMemcpyAsync(&n, c->n_rows, sizeof(int), cudaMemcpyDeviceToHost, st));
        GCP_CUDA_CHECK(cudaStreamSynchronize(st));
        c->prof_tail_images += n;
    }
    return 0;
}

// Inverse model on consecutive rows of the zero-padded latent sequence seq [B][200][128] and state regressor on every
// row (InverseModel.full_seq_forward, inverse_mdl.py:110-134; base_gcp.py:252-256).
static int run_pair_heads(gcpb200_ctx* c, cudaStream_t st, const long long* end_ind, int B, float* actions,
                          float* regressed_state) {
    // c->seqb holds the sequences as bf16 rows (cand, t), c->max_len + 1 rows per candidate (gather_frames_kernel): the
    // pair [frame t | frame t + 1] is two row-shifted K segments of the same array, no pair matrix is materialised.
    const LevelGeom flat = {(B + 127) / 128 * 128, 0, c->depth};
    const int L1 = c->max_len + 1;
    const int rows = (B * L1 + 127) / 128 * 128;
    (void)end_ind;
    if ((actions && !c->has_inv) || (regressed_state && !c->has_state)) {
        gcp_set_error("actions / regressed_state requested but the loaded state dict has no inv_mdl / state_regressor weights");
        return -1;
    }
    if (rows > c->pair_rows) {
        gcp_set_error("run_pair_heads: %d rows exceed the scratch capacity %d", rows, c->pair_rows);
        return -1;
    }
    if (actions) {
        CHECK(mlp_body(c, st, c->inv_mdl, rows, flat, {seg(c->seqb, 0, NZ_ENC, ROW_LEVEL, 0), seg(c->seqb, 0, NZ_ENC, ROW_LEVEL, 1)}));
        CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->inv_mdl.mid_k)}, c->inv_mdl.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 2, 2)));
        GCP_CUDA_CHECK(cudaMemcpy2DAsync(actions, (size_t)c->max_len * 2 * 4, c->rowcost, (size_t)L1 * 2 * 4, (size_t)c->max_len * 2 * 4, B,
                                         cudaMemcpyDeviceToDevice, st));
    }
    if (regressed_state) {
        CHECK(mlp_body(c, st, c->state_reg, rows, flat, {seg(c->seqb, 0, NZ_ENC)}));
        CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->state_reg.mid_k)}, c->state_reg.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 2, 2)));
        GCP_CUDA_CHECK(cudaMemcpy2DAsync(regressed_state, (size_t)c->max_len * 2 * 4, c->rowcost, (size_t)L1 * 2 * 4,
                                         (size_t)c->max_len * 2 * 4, B, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// Posterior inputs of the training-phase tree (null = prior rollout)
struct PosteriorArgs {
    const float* inf_seq;   // [B][200][128] inference encoding of every frame
    const int* tstep;       // [B][255] matched frame of every node (depth-first)
    float *q_mu, *q_ls;     // [B][255][256]
};

// One level of SubgoalTreeLayer.produce_tree (gcp/prediction/utils/tree_utils.py:21-44) = TreeModule.produce_subgoal on all
// B * 2^l nodes of level l (tree_module.py:67-114): prior (+ posterior), reparametrisation, TreeLSTM, output latent.
static int tree_level(gcpb200_ctx* c, cudaStream_t st, int l, int B, int Bp, const float* z, float* mu_df, float* ls_df,
                      const PosteriorArgs* post, float* e_df = nullptr, bool pruned = false) {
    const int goal_row0 = (c->n_nodes + 1) * Bp;
    {
        const LevelW& L = c->lvl[c->tied ? 0 : l];
        // planner mode: only the (node, candidate tile) pairs of this level's work list; the launches are shaped for the whole
        // level and read the live row count (c->tree_rows[l]) on the device
        const LevelGeom g = {Bp, l, c->depth, pruned ? c->tree_tiles + tree_tiles_offset(l, Bp >> 7) : nullptr};
        const GemmDyn dynv = {c->tree_rows + l, 0};
        const GemmDyn* dyn = pruned ? &dynv : nullptr;
        const int rows = Bp << l;
        // split-linear projections of the parents' LSTM state
        auto project = [&](cudaStream_t s) -> int {
            Seg a = seg(c->hid, 0, HID, ROW_LEFT), b = seg(c->hid, 0, HID, ROW_RIGHT);
            const int gc[6] = {0, 2 * HID, 4 * HID, HID, 3 * HID, 5 * HID};
            a.group_cols = b.group_cols = HID;
            for (int q = 0; q < 6; ++q) a.group_col[q] = b.group_col[q] = gc[q];
            // Both halves are stored as bf16.  Measured (profiles/r2i_fp32_cell_state.txt): keeping the projected cell state
            // in fp32 (split_col = 3 * HID -> an fp32 array read by the LSTM epilogue) leaves the latent error where it is
            // (max-rel 7.5e-3 either way: it is the bf16 rounding of the GEMM OPERANDS, which the SIMT cross-check kernels
            // with fp32 accumulation reproduce) and costs 1.2 ms per 1024-candidate rollout in extra HBM traffic.
            EpiParams e = epi_linear(ACT_NONE, c->sh.p, 6 * HID, nullptr, 0, 6 * HID);
            return gemm(c, s, rows, g, {a, b}, L.proj, 256, EPI_LINEAR, e, 0, -1, dyn);
        };
        // A level whose row tiles do not fill the GPU is a chain of latency-bound launches: its projection GEMM (which needs
        // only the parents' state, complete since the previous level) runs beside the prior -> reparametrisation -> embed
        // chain on a side stream.  Not at level 0 (the parents' state comes from the initializer below) and not in the
        // training phase (one more chain, the posterior, shares the buffers in a different order).
        const bool proj_aside = l > 0 && post == nullptr && (rows >> 7) <= c->sms && !c->use_ref;
        if (proj_aside) {
            GCP_CUDA_CHECK(cudaEventRecord(c->ev_proj_fork, st));
            GCP_CUDA_CHECK(cudaStreamWaitEvent(c->proj_stream, c->ev_proj_fork, 0));
            CHECK(project(c->proj_stream));
            GCP_CUDA_CHECK(cudaEventRecord(c->ev_proj_join, c->proj_stream));
        }
        // prior p(z | e_l, e_r) and reparametrisation
        const std::vector<Seg> par = {seg(c->lat, 0, NZ_ENC, ROW_LEFT), seg(c->lat, 0, NZ_ENC, ROW_RIGHT)};
        CHECK(mlp_body(c, st, L.prior, rows, g, par, dyn));
        {
            EpiParams e;
            memset(&e, 0, sizeof(e));
            e.z = z; e.n_cand = B; e.nz = NZ_VAE;
            e.out_bf16 = c->zeta.p; e.out_bf16_ld = NZ_VAE;
            e.mu_out = mu_df; e.ls_out = ls_df;
            CHECK(gemm(c, st, rows, g, {seg(c->tb, 0, L.prior.mid_k)}, L.prior.head, 256, EPI_REPARAM, e, 0, -1, dyn));
        }
        if (post != nullptr) {
            // training phase: z ~ q(z | e_l, e_r, e_tilde), e_tilde = inference encoding of the frame the node is matched
            // to (tree_module.py:86-95, tree/inference.py:16-36); the sample overwrites the prior's in c->zeta
            const size_t n = (size_t)rows * 128;
            gather_etilde_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(post->inf_seq, post->tstep, g, B, c->max_len, c->xb.p, HID);
            LAUNCH_CHECK();
            CHECK(mlp_body(c, st, L.q, rows, g, {par[0], par[1], seg(c->xb, 0, NZ_ENC)}));
            EpiParams e;
            memset(&e, 0, sizeof(e));
            e.z = z; e.n_cand = B; e.nz = NZ_VAE;
            e.out_bf16 = c->zeta.p; e.out_bf16_ld = NZ_VAE;
            e.mu_out = post->q_mu; e.ls_out = post->q_ls;
            CHECK(gemm(c, st, rows, g, {seg(c->tb, 0, L.q.mid_k)}, L.q.head, 256, EPI_REPARAM, e));
        }
        const std::vector<Seg> par_z = {par[0], par[1], seg(c->zeta, 0, NZ_VAE)};
        if (l == 0) {
            // MLPLSTMCellInitializer: hidden states of the two root parents (slots 0 and 256)
            CHECK(mlp_body(c, st, L.init, rows, g, par_z, dyn));
            CHECK(gemm(c, st, rows, g, {seg(c->tb, 0, L.init.mid_k)}, L.init.head, 256, EPI_LINEAR,
                       epi_linear(ACT_NONE, c->hid.p, STATE, nullptr, 0, STATE), 0, -1, dyn));
            CHECK(gemm(c, st, rows, g, {seg(c->tb, 0, L.init.mid_k)}, L.init_head_r, 256, EPI_LINEAR,
                       epi_linear(ACT_NONE, c->hid.p + (size_t)goal_row0 * STATE, STATE, nullptr, 0, STATE), 0, -1, dyn));
        }
        if (!proj_aside) CHECK(project(st));
        // embed
        {
            EpiParams e = epi_linear(ACT_NONE, c->xa.p, HID, nullptr, 0, HID);
            const std::vector<Seg> in = {par[0], par[1], seg(c->zeta, 0, NZ_VAE), seg(c->lat, 0, NZ_ENC, ROW_CAND, 0),
                                         seg(c->lat, 0, NZ_ENC, ROW_CAND, goal_row0)};
            CHECK(gemm(c, st, rows, g, in, L.embed_main, 256, EPI_LINEAR, e, 0, -1, dyn));
        }
        if (proj_aside) GCP_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_proj_join, 0));
        // three LSTM cells; the new (h, c) of every non-leaf node goes to the slot-major state array
        DevBuf* xin = &c->xa;
        DevBuf* xout = &c->xb;
        for (int i = 0; i < N_LSTM; ++i) {
            EpiParams e;
            memset(&e, 0, sizeof(e));
            e.c_prev = c->sh.p; e.c_prev_ld = 6 * HID; e.c_prev_col0 = (3 + i) * HID;
            e.out_bf16 = xout->p; e.out_bf16_ld = HID;
            e.hid = c->hid.p; e.hid_ld = STATE; e.hid_col0 = 2 * HID * i; e.hidden = HID;
            e.write_hid = (l < c->depth - 1);
            CHECK(gemm(c, st, rows, g, {seg(*xin, 0, HID), seg(c->sh, i * HID, HID)}, L.lstm[i], 256, EPI_LSTM, e, 0, -1, dyn));
            std::swap(xin, xout);
        }
        // output linear -> node latent e' (raw, no activation) at the node's slot
        // (fp32 copy: slot-major, or -- e_df given -- straight into the caller's depth-first [B][255][128] output)
        EpiParams eo = epi_linear(ACT_NONE, c->lat.p, NZ_ENC, c->lat_f32, NZ_ENC, NZ_ENC, ROW_SELF, ROW_SELF);
        if (e_df != nullptr) {
            eo.out_f32 = e_df;
            eo.out_f32_df = c->n_nodes;
            eo.n_cand = B;
        }
        CHECK(gemm(c, st, rows, g, {seg(*xin, 0, HID)}, L.out, 128, EPI_LINEAR, eo, 0, -1, dyn));
    }

    return 0;
}

extern "C" int gcpb200_rollout(gcpb200_ctx* c, const gcpb200_rollout_io* io, void* stream) {
    if (!io) {
        gcp_set_error("null io");
        return -1;
    }
    CHECK(check_ready(c, io->B));
    if (c->model != GCPB200_MODEL_TREE && c->model != GCPB200_MODEL_TREE_ADAPTIVE) {
        gcp_set_error("gcpb200_rollout needs a context created with model = GCPB200_MODEL_TREE or _TREE_ADAPTIVE");
        return -1;
    }
    const bool adaptive = c->model == GCPB200_MODEL_TREE_ADAPTIVE;
    if (adaptive && (io->existence || io->model_enc_seq || io->actions || io->regressed_state)) {
        gcp_set_error("the adaptive model has no existence predictor / balanced pruning heads (use distances, pruned_nodes)");
        return -1;
    }
    if (!adaptive && (io->distances || io->pruned_nodes || io->pruned_len)) {
        gcp_set_error("distances / pruned_nodes need a context created with model = GCPB200_MODEL_TREE_ADAPTIVE");
        return -1;
    }
    if (!io->I_0 || !io->I_g || !io->z) {
        gcp_set_error("gcpb200_rollout: I_0, I_g and z are required");
        return -1;
    }
    const bool fused_l2 = io->l2_cost != nullptr;
    if ((io->decode_kept_only || fused_l2) && (adaptive || c->use_ref)) {
        gcp_set_error("decode_kept_only / l2_cost need the balanced GCPB200_MODEL_TREE model");
        return -1;
    }
    if (fused_l2 && !io->l2_goal) {
        gcp_set_error("l2_cost needs l2_goal");
        return -1;
    }
    const bool kept_only = io->decode_kept_only != 0 && (io->images_df != nullptr || fused_l2);
    if (io->tree_kept_only && (!io->decode_kept_only || io->existence)) {
        gcp_set_error("tree_kept_only needs decode_kept_only and no existence output (pruned-away nodes are not computed)");
        return -1;
    }
    if (io->tree_kept_only && ((io->B + 127) >> 7) > TREE_TPN_MAX) {
        gcp_set_error("tree_kept_only: at most %d candidates per call", TREE_TPN_MAX * 128);
        return -1;
    }
    const bool tree_pruned = io->tree_kept_only != 0 && !adaptive && !c->use_ref;
    const bool decode_all = !kept_only && (io->images_df != nullptr || fused_l2);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = io->B, Bp = (B + 127) / 128 * 128;
    const LevelGeom flat = {Bp, 0, c->depth};
    const int goal_row0 = (c->n_nodes + 1) * Bp;

    ProfScope total_scope(c, st, 5);
    if (io->z_host) {
        // Noise upload, overlapped: the copy stream gathers the rows of levels 0-3, then levels 4, 5, 6, 7 one by one
        // from pinned host memory; the rollout stream waits for each set just before the level that consumes it.
        trace_mark(st, "start");
        GCP_CUDA_CHECK(cudaEventRecord(c->ev_copy_start, st));       // earlier users of the staging buffer are done
        GCP_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->ev_copy_start, 0));
        // node = a*k + b, k < cnt: set 0 = levels 0 .. depth-5 together, sets 1-4 = levels depth-4 .. depth-1 one by one
        // (depth 8: {16,15,15}, {16,7,16}, {8,3,32}, {4,1,64}, {2,0,128}); trees of fewer than 5 levels upload everything at once
        int sets[5][3];
        const int D = c->depth, n_sets = D >= 5 ? 5 : 1;
        if (D >= 5) {
            sets[0][0] = 16; sets[0][1] = 15; sets[0][2] = (1 << (D - 4)) - 1;
            for (int i = 1; i < 5; ++i) {
                const int l = D - 5 + i;
                sets[i][0] = 1 << (D - l); sets[i][1] = (1 << (D - 1 - l)) - 1; sets[i][2] = 1 << l;
            }
        } else {
            sets[0][0] = 1; sets[0][1] = 0; sets[0][2] = c->n_nodes;
        }
        // Launch shape (measured with GCPB200_TRACE=1, profiles/r1j_upload_interference.txt): PCIe saturates at about
        // 8 k outstanding 16-byte reads.  Reads beyond that queue inside the GPU's memory system and slow the tensor-core
        // kernels of the rollout stream (12-16 k outstanding: GEMMs and decoder 1.5x slower; 19 k: they stall until the
        // upload ends), fewer leave PCIe idle.  64 blocks x 32 threads x 4 reads in flight = 8192.
#if GCP_VERIFY
        static int up_grid = 0, up_block = 0;
        if (up_grid == 0) {
            const char* eg = getenv("GCPB200_UPLOAD_GRID");
            const char* eb = getenv("GCPB200_UPLOAD_BLOCK");
            up_grid = eg ? atoi(eg) : 64;
            up_block = eb ? atoi(eb) : 32;
            if (up_grid < 1 || up_block < 32 || up_block > 128 || (up_block & 31)) {
                up_grid = 64;
                up_block = 32;
            }
        }
#else
        const int up_grid = 64, up_block = 32;
#endif
        for (int i = 0; i < n_sets; ++i) {
            upload_rows_kernel<<<up_grid, up_block, 0, c->copy_stream>>>(reinterpret_cast<const float4*>(io->z_host),
                                                                      reinterpret_cast<float4*>(const_cast<float*>(io->z)), B,
                                                                      c->n_nodes, NZ_VAE / 4, sets[i][0], sets[i][1], sets[i][2]);
            LAUNCH_CHECK();
            GCP_CUDA_CHECK(cudaEventRecord(c->ev_copy[i], c->copy_stream));
            static const char* names[5] = {"up_l0-3", "up_l4", "up_l5", "up_l6", "up_l7"};
            trace_mark(c->copy_stream, names[i]);
        }
    }
    ProfScope* scope = new ProfScope(c, st, 0);
    {
        CommonIO cio = {io->I_0, io->I_g, io->images_shared, io->end_ind, io->seed, B, io->e_0, io->e_g, io->seq_len_logits,
                        io->end_ind_out, io->sort_sampled_lengths};
        CHECK(run_encoder_length(c, st, cio, Bp, goal_row0));
    }
    const std::vector<Seg> ctx_in = {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, 0), seg(c->lat, 0, NZ_ENC, ROW_LEVEL, goal_row0)};

    delete scope;
    // ---- 2b. beside the tree recursion: everything that needs only the encoder skips and the rollout length
    const bool need_map = kept_only || fused_l2 || io->model_enc_seq || io->actions || io->regressed_state;
    const bool decodes = kept_only || decode_all;
    bool prep_pending = need_map || decodes;
    if (prep_pending) {
        cudaStream_t ps = c->prep_stream;
        GCP_CUDA_CHECK(cudaEventRecord(c->ev_prep_fork, st));
        GCP_CUDA_CHECK(cudaStreamWaitEvent(ps, c->ev_prep_fork, 0));
        if (need_map) CHECK(compute_frame_map(c, c->end_ind, B, ps));
        if (kept_only) {
            kept_offsets_kernel<<<1, 1024, 0, ps>>>(c->end_ind, B, 1, c->row_off, c->n_rows);
            LAUNCH_CHECK();
        }
        if (decodes) CHECK(decoder_prepare(c, ps, io->images_shared, B, Bp));
        GCP_CUDA_CHECK(cudaEventRecord(c->ev_prep_join, ps));
    }
    auto prep_join = [&]() -> int {
        if (prep_pending) GCP_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_prep_join, 0));
        prep_pending = false;
        return 0;
    };
    scope = new ProfScope(c, st, 1);
    // ---- 3. tree recursion, level by level (SubgoalTreeLayer.produce_tree)
    // Level-ordered decoding: a node can be decoded as soon as its level is done, so the decoder runs in three parts --
    // levels 0-5 (63 nodes) before tree level 6, level 6 (64 nodes) before tree level 7, level 7 (128 nodes) after it.
    // Same work in total; with host-resident noise it puts 2 ms of tensor-bound work in front of each of the two big
    // uploads (level 6: 67 MB, level 7: 133 MB at 1024 candidates) instead of stalling the recursion on PCIe.
    const bool level_ordered = decode_all && !c->use_ref && c->depth >= 3;
    const float* l2_goal = fused_l2 ? io->l2_goal : nullptr;
    float* e_df = io->e_df ? io->e_df : c->e_df;
    for (int l = 0; l < c->depth; ++l) {
        if (level_ordered && l >= c->depth - 2) {
            trace_mark(st, l == c->depth - 2 ? "tree_l5_end" : "tree_l6_end");
            delete scope;
            scope = nullptr;
            if (l == c->depth - 2) {
                CHECK(prep_join());
                CHECK(decoder_slots(c, st, io->images_shared, B, Bp, 4, 4, (1 << (c->depth - 2)) - 1, io->images_df, c->n_nodes, io->I_0, io->I_g, l2_goal));
            } else {
                CHECK(decoder_slots(c, st, io->images_shared, B, Bp, 2, 4, 1 << (c->depth - 2), io->images_df, c->n_nodes, io->I_0, io->I_g, l2_goal));
            }
            scope = new ProfScope(c, st, 1);
            trace_mark(st, l == c->depth - 2 ? "dec_l0-5_end" : "dec_l6_end");
        }
        if (io->z_host && (l == 0 || (c->depth >= 5 && l >= c->depth - 4)))
            GCP_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_copy[l == 0 ? 0 : l - (c->depth - 5)], 0));
        if ((io->mu_df == nullptr) != (io->log_sigma_df == nullptr)) {
            gcp_set_error("mu_df and log_sigma_df must be given together");
            return -1;
        }
        if (tree_pruned && l == 0) {
            // the rollout length is known (sampled / injected in run_encoder_length): work lists of every level
            tree_worklists_kernel<<<c->depth, 1024, 0, st>>>(c->end_ind, B, Bp, 1, c->tree_tiles, c->tree_rows);
            LAUNCH_CHECK();
        }
        CHECK(tree_level(c, st, l, B, Bp, io->z, io->mu_df, io->log_sigma_df, nullptr, e_df, tree_pruned));
    }

    trace_mark(st, "tree_l7_end");
    delete scope;
    scope = new ProfScope(c, st, 4);
    // ---- 4. existence predictor (the depth-first fp32 latents were written by the output GEMM of every level)
    if (io->existence) {
        const int rows = c->n_nodes * Bp;
        CHECK(mlp_body(c, st, c->existence, rows, flat, {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, Bp)}));
        CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->existence.mid_k)}, c->existence.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->exist_slot + Bp, 1, 1)));
        const size_t n = (size_t)B * c->n_nodes;
        slot_to_df_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->exist_slot, Bp, B, c->n_nodes, 1, 1, io->existence);
        LAUNCH_CHECK();
    }

    delete scope;
    scope = nullptr;
    // ---- 5. decoder over all 255 node latents
    CHECK(prep_join());
    if (level_ordered)      // level 7 = the odd slots
        CHECK(decoder_slots(c, st, io->images_shared, B, Bp, 1, 2, c->n_nodes / 2 + 1, io->images_df, c->n_nodes, io->I_0, io->I_g, l2_goal));
    else if (kept_only)
        CHECK(decoder_kept(c, st, io->images_shared, B, Bp, io->images_df, l2_goal));
    else if (io->images_df)    // (decoder_prepare ran on the side stream)
        CHECK(decoder_slots(c, st, io->images_shared, B, Bp, 1, 1, c->n_nodes, io->images_df, c->n_nodes, io->I_0, io->I_g));
    if (fused_l2) {
        cost_from_frames_kernel<<<B, 32, 0, st>>>(c->frame_sq, kept_only ? nullptr : c->frame_node, c->row_off, c->end_ind, c->n_nodes,
                                                  c->max_len, io->l2_dense, io->l2_final_step_weight, 1, io->l2_cost);
        LAUNCH_CHECK();
    }
    if (adaptive && (io->distances || io->pruned_nodes || io->pruned_len)) {
        // AdaptiveBinding.prune_sequence: distance predictor on consecutive depth-first latents, then compaction
        ProfScope dsc(c, st, 4);
        if (!io->pruned_nodes || !io->pruned_len) {
            gcp_set_error("pruned_nodes and pruned_len must be given together");
            return -1;
        }
        const int rows = (B * c->n_nodes + 127) / 128 * 128;
        make_pairs_kernel<<<(unsigned)(((size_t)rows * 256 + 255) / 256), 256, 0, st>>>(e_df, c->end_ind, nullptr, B, c->n_nodes, rows,
                                                                                       c->pairs.p);
        LAUNCH_CHECK();
        CHECK(mlp_body(c, st, c->distance_pred, rows, flat, {seg(c->pairs, 0, 256)}));
        CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->distance_pred.mid_k)}, c->distance_pred.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 1, 1)));
        if (io->distances)
            GCP_CUDA_CHECK(cudaMemcpy2DAsync(io->distances, (c->n_nodes - 1) * 4, c->rowcost, c->n_nodes * 4, (c->n_nodes - 1) * 4, B,
                                             cudaMemcpyDeviceToDevice, st));
        const float thr = io->prune_threshold > 0.f ? io->prune_threshold : 0.5f;
        adaptive_prune_kernel<<<B, 256, 0, st>>>(c->rowcost, c->n_nodes, c->n_nodes, logf(thr / (1.0f - thr)), io->pruned_nodes,
                                                 io->pruned_len, nullptr);
        LAUNCH_CHECK();
    }

    ProfScope asc(c, st, 4);
    // ---- 6. pruned latent sequence + inverse model + state regressor (run_auxilliary_models)
    if (io->model_enc_seq || io->actions || io->regressed_state) {
        float* seq = io->model_enc_seq ? io->model_enc_seq : c->seq;
        const size_t n = (size_t)B * c->max_len * (NZ_ENC / 4);
        gather_frames_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e_df, c->frame_node, c->end_ind, B, c->n_nodes, c->max_len,
                                                                           NZ_ENC / 4, seq, c->seqb.p);
        LAUNCH_CHECK();
        if (io->actions || io->regressed_state) CHECK(run_pair_heads(c, st, c->end_ind, B, io->actions, io->regressed_state));
    }
    trace_mark(st, "end");
    trace_dump();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Training-phase forward + loss (BASELINE config 1)
// ---------------------------------------------------------------------------------------------
static int ensure_train_ws(gcpb200_ctx* c) {
    gcpb200_ctx::TrainWS& w = c->tw;
    if (w.Bcap > 0) return 0;
    const size_t Bc = 128, n_img = Bc * (MAX_LEN + 2), rows = (size_t)N_NODES * Bc;
    int rc = 0;
    rc |= dalloc(c, &w.y1, n_img * 2048, false);
    rc |= dalloc(c, &w.y2, n_img * 1024, false);
    rc |= dalloc(c, &w.enc_seq, Bc * MAX_LEN * 128);
    rc |= dalloc(c, &w.h1, Bc * MAX_LEN * 128);
    rc |= dalloc(c, &w.h2, Bc * MAX_LEN * 128);
    rc |= dalloc(c, &w.inf_seq, Bc * MAX_LEN * 128);
    rc |= dalloc(c, &w.stats, 4096);
    rc |= dalloc(c, &w.tstep, Bc * N_NODES);
    rc |= dalloc(c, &w.keep, Bc * N_NODES);
    rc |= make_buf(c, &w.x1, rows, 1024);
    rc |= make_buf(c, &w.x2, rows, 2048);
    rc |= make_buf(c, &w.x3, rows, 4096);
    for (int i = 0; i < 4; ++i) rc |= dalloc(c, &w.pq[i], Bc * N_NODES * NZ_VAE);
    rc |= dalloc(c, &w.raw_mu, Bc * 64 * 1024 * 16, false);
    rc |= dalloc(c, &w.raw_ls, Bc * 64 * 1024 * 16, false);
    rc |= dalloc(c, &w.nll_bt, Bc * MAX_LEN);
    rc |= dalloc(c, &w.kl_b, Bc);
    rc |= dalloc(c, &w.reg, Bc * MAX_LEN * 2 + 512);
    rc |= dalloc(c, &w.inv_pred, 128 * 2);
    rc |= dalloc(c, &w.cost_pred, 128);
    rc |= dalloc(c, &w.exist_df, Bc * N_NODES);
    rc |= dalloc(c, &w.cost_tgt, Bc);
    if (rc) return -1;
    cudaError_t e = cudaSuccess;
#if GCP_VERIFY
    e = cudaFuncSetAttribute(dec_tail_nll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES);
#endif
    if (e != cudaSuccess) {
        gcp_set_error("cudaFuncSetAttribute(dec_tail_nll_kernel) failed: %s", cudaGetErrorString(e));
        return -1;
    }
    w.Bcap = (int)Bc;
    return 0;
}

// One decoder layer's batch-statistic BatchNorm + ReLU, in place on the raw bf16 GEMM output of all 255 * Bp node rows
static int bn_layer(gcpb200_ctx* c, cudaStream_t st, DevBuf& x, int kind, int B, int Bp, double* stats, int bn_idx, int C,
                    int spatial) {
    const int rows = N_NODES * Bp;
    bn_stats_kernel<<<dim3(x.ld / 256, (rows + 63) / 64), 256, 0, st>>>(x.p, rows, x.ld, Bp, B, kind, stats);
    LAUNCH_CHECK();
    const size_t n_vec = (size_t)rows * x.ld / 8;
    bn_apply_relu_kernel<<<c->sms * 8, 256, 0, st>>>(x.p, n_vec, x.ld, kind, stats, (double)B * N_NODES * spatial, c->bn_g[bn_idx],
                                                     c->bn_b[bn_idx], C);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_forward_loss(gcpb200_ctx* c, const gcpb200_train_io* io, void* stream) {
    if (!io) {
        gcp_set_error("null io");
        return -1;
    }
    CHECK(check_ready(c, io->B));
    if (c->model != GCPB200_MODEL_TREE || !c->has_train || !c->has_cost || !c->has_inv || !c->has_state) {
        gcp_set_error("gcpb200_forward_loss needs a GCPB200_MODEL_TREE context with attach_cost_mdl = 1 and a state dict that "
                      "holds the training tensors (inf_encoder.*, inference.q.*, inv_mdl.*, state_regressor.*, cost_mdl.*)");
        return -1;
    }
    if (io->B > 128) {
        gcp_set_error("gcpb200_forward_loss: B = %d > 128 sequences per call", io->B);
        return -1;
    }
    if (!io->traj_seq || !io->pad_mask || !io->end_ind || !io->I_0 || !io->I_g || !io->states || !io->actions || !io->eps ||
        !io->inv_t0 || !io->inv_t1 || !io->cost_start || !io->cost_end || !io->losses) {
        gcp_set_error("gcpb200_forward_loss: every input pointer except cost_target, and `losses`, are required");
        return -1;
    }
    CHECK(ensure_train_ws(c));
    gcpb200_ctx::TrainWS& w = c->tw;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = io->B, Bp = 128, T = MAX_LEN;
    const LevelGeom flat = {Bp, 0, DEPTH};
    const int goal_row0 = 256 * Bp;
    double* st1 = w.stats;                 // [3][32][2]
    double* st2 = st1 + 3 * 32 * 2;        // [3][64][2]
    double* gn = st2 + 3 * 64 * 2;         // [128][8][2]
    double* sd1 = gn + 128 * 8 * 2;        // [64][2]
    double* sd2 = sd1 + 64 * 2;            // [64][2] (32 used)
    double* sd3 = sd2 + 64 * 2;            // [64][2] (16 used)
    GCP_CUDA_CHECK(cudaMemsetAsync(w.stats, 0, 4096 * sizeof(double), st));
    GCP_CUDA_CHECK(cudaMemcpyAsync(c->end_ind, io->end_ind, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));

    // ---- 1. encoders with batch statistics: all frames, start images, goal images (base_gcp.py:184-209)
    {
        EncTrainArgs a;
        memset(&a, 0, sizeof(a));
        a.img[0] = io->traj_seq; a.img[1] = io->I_0; a.img[2] = io->I_g;
        a.n[0] = B * T; a.n[1] = B; a.n[2] = B;
        a.W = c->enc;
        a.g1 = c->bn_g[0]; a.b1 = c->bn_b[0]; a.g2 = c->bn_g[1]; a.b2 = c->bn_b[1];
        a.y1 = w.y1; a.y2 = w.y2; a.st1 = st1; a.st2 = st2;
        a.enc_seq = w.enc_seq; a.lat_f32 = c->lat_f32; a.lat_bf16 = c->lat.p; a.row0_a = 0; a.row0_b = goal_row0;
        a.skip0 = c->s0; a.skip2 = c->s2; a.skip2_bf16 = c->s2b.p;
        a.w1t = c->enc_w1t; a.w2t = c->enc_w2t; a.w3t = c->enc_w3t;
        const int n_img = B * (T + 2);
        const int n_quads = (n_img + ENCB_IMGS - 1) / ENCB_IMGS, n_blk = (n_img + ENCC_IMGS - 1) / ENCC_IMGS;
        // persistent over images: 3 / 2 / 3 CTAs per SM fit the shared memory of the three passes
        enc_train_a_kernel<<<std::min(n_img, 3 * c->sms), ENC_T, ENCA_SMEM, st>>>(a);
        LAUNCH_CHECK();
        enc_train_b_kernel<<<std::min(n_quads, 2 * c->sms), ENC_T, ENCB_SMEM, st>>>(a);
        LAUNCH_CHECK();
        enc_train_c_kernel<<<std::min(n_blk, 3 * c->sms), ENC_T, ENCC_SMEM, st>>>(a);
        LAUNCH_CHECK();
    }
    if (io->e_0) GCP_CUDA_CHECK(cudaMemcpyAsync(io->e_0, c->lat_f32, (size_t)B * NZ_ENC * 4, cudaMemcpyDeviceToDevice, st));
    if (io->e_g)
        GCP_CUDA_CHECK(cudaMemcpyAsync(io->e_g, c->lat_f32 + (size_t)goal_row0 * NZ_ENC, (size_t)B * NZ_ENC * 4,
                                       cudaMemcpyDeviceToDevice, st));
    if (io->enc_traj_seq)
        GCP_CUDA_CHECK(cudaMemcpyAsync(io->enc_traj_seq, w.enc_seq, (size_t)B * T * 128 * 4, cudaMemcpyDeviceToDevice, st));
    // ---- 2. inference encoder: three conv1d over time (subnetworks.py:120-147)
    {
        const dim3 grid((T + 7) / 8, B);
        conv1d_k3_kernel<<<grid, 128, 0, st>>>(w.enc_seq, 128, 1, c->ie_w[0], c->ie_b[0], nullptr, nullptr, nullptr, ACT_LRELU, w.h1,
                                               nullptr, T);
        LAUNCH_CHECK();
        conv1d_k3_kernel<<<grid, 128, 0, st>>>(w.h1, 128, 0, c->ie_w[1], nullptr, nullptr, nullptr, nullptr, ACT_NONE, w.h2, gn, T);
        LAUNCH_CHECK();
        conv1d_k3_kernel<<<grid, 128, 0, st>>>(w.h2, 128, 0, c->ie_w[2], c->ie_b[2], gn, c->ie_gn_g, c->ie_gn_b, ACT_NONE, w.inf_seq,
                                               nullptr, T);
        LAUNCH_CHECK();
    }
    if (io->inf_enc_seq)
        GCP_CUDA_CHECK(cudaMemcpyAsync(io->inf_enc_seq, w.inf_seq, (size_t)B * T * 128 * 4, cudaMemcpyDeviceToDevice, st));
    // ---- 3. length predictor logits (misc.py:38-51)
    {
        const std::vector<Seg> ctx_in = {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, 0), seg(c->lat, 0, NZ_ENC, ROW_LEVEL, goal_row0)};
        CHECK(mlp_body(c, st, c->length_pred, Bp, flat, ctx_in));
        CHECK(gemm(c, st, Bp, flat, {seg(c->tb, 0, c->length_pred.mid_k)}, c->length_pred.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->logits, 256, MAX_LEN)));
        if (io->seq_len_logits)
            GCP_CUDA_CHECK(cudaMemcpy2DAsync(io->seq_len_logits, MAX_LEN * 4, c->logits, 256 * 4, MAX_LEN * 4, B,
                                             cudaMemcpyDeviceToDevice, st));
    }
    // ---- 4. frame <-> node matching (integer)
    GCP_CUDA_CHECK(cudaMemsetAsync(c->frame_node, 0, (size_t)B * MAX_LEN * sizeof(int), st));
    match_tables_kernel<<<(B * N_NODES + 255) / 256, 256, 0, st>>>(c->end_ind, B, DEPTH, MAX_LEN, c->frame_node, w.tstep, w.keep);
    LAUNCH_CHECK();
    if (io->match_timesteps)
        GCP_CUDA_CHECK(cudaMemcpyAsync(io->match_timesteps, w.tstep, (size_t)B * N_NODES * 4, cudaMemcpyDeviceToDevice, st));
    // ---- 5. tree with the approximate posterior
    float* p_mu = io->p_mu ? io->p_mu : w.pq[0];
    float* p_ls = io->p_log_sigma ? io->p_log_sigma : w.pq[1];
    float* q_mu = io->q_mu ? io->q_mu : w.pq[2];
    float* q_ls = io->q_log_sigma ? io->q_log_sigma : w.pq[3];
    {
        PosteriorArgs post = {w.inf_seq, w.tstep, q_mu, q_ls};
        for (int l = 0; l < DEPTH; ++l) CHECK(tree_level(c, st, l, B, Bp, io->eps, p_mu, p_ls, &post));
    }
    float* e_df = io->e_df ? io->e_df : c->e_df;
    {
        const size_t n = (size_t)B * N_NODES * NZ_ENC;
        slot_to_df_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->lat_f32, Bp, B, N_NODES, NZ_ENC, NZ_ENC, e_df);
        LAUNCH_CHECK();
    }
    float* exist = io->existence ? io->existence : w.exist_df;
    {
        const int rows = N_NODES * Bp;
        CHECK(mlp_body(c, st, c->existence, rows, flat, {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, Bp)}));
        CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->existence.mid_k)}, c->existence.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->exist_slot + Bp, 1, 1)));
        const size_t n = (size_t)B * N_NODES;
        slot_to_df_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->exist_slot, Bp, B, N_NODES, 1, 1, exist);
        LAUNCH_CHECK();
    }
    // ---- 6. decoder over all nodes, BatchNorm with the statistics of this batch (tree_dense_rec.py:41-44)
    {
        const int rows = N_NODES * Bp;
        skip_prep_kernel<<<B, 256, 0, st>>>(c->s0, c->skip_up, B);
        LAUNCH_CHECK();
        CHECK(gemm(c, st, Bp, flat, {seg(c->s2b, 0, 1024)}, c->dec2st, 256, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->rowbias2, 2048, 2048)));
        CHECK(gemm(c, st, rows, flat, {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, Bp)}, c->dec1t, 256, EPI_LINEAR,
                   epi_linear(ACT_NONE, w.x1.p, 1024, nullptr, 0, 1024)));
        CHECK(bn_layer(c, st, w.x1, 1, B, Bp, sd1, 2, 64, 16));
        {
            EpiParams e = epi_linear(ACT_NONE, w.x2.p, 2048, nullptr, 0, 2048);
            e.rowbias = c->rowbias2;
            e.rowbias_ld = 2048;
            CHECK(gemm(c, st, rows, flat, {seg(w.x1, 0, 1024)}, c->dec2xt, 256, EPI_LINEAR, e));
        }
        CHECK(bn_layer(c, st, w.x2, 2, B, Bp, sd2, 3, 32, 64));
        {
            Seg a3 = seg(w.x2, 0, 1024);
            a3.group_cols = 256;
            for (int q = 0; q < 16; ++q) a3.group_col[q] = dec3_window_row0(q & 7) * 256;
            CHECK(gemm(c, st, rows, flat, {a3}, c->dec3t, 256, EPI_LINEAR, epi_linear(ACT_NONE, w.x3.p, 4096, nullptr, 0, 4096)));
        }
        CHECK(bn_layer(c, st, w.x3, 3, B, Bp, sd3, 4, 16, 256));
        // skip half of the 32->16 tail conv per sequence, quad layout of the tcgen05 tail kernel
        skip_term3_kernel<<<dim3(8, B), 128, 0, st>>>(c->skip_up, c->w4p, c->b4, c->s4);
        LAUNCH_CHECK();
        if (io->images_df) {
            // DLM mean image of every node with the rollout's tcgen05 tail kernel (tree.df.images; logging only)
            DecTail3Args a;
            memset(&a, 0, sizeof(a));
            a.x3 = w.x3.p; a.s4 = c->s4; a.s4_stride = 256 * 64;
            a.w4 = c->z4; a.w5 = c->z5; a.b5h = c->b5h;
            a.images = io->images_df; a.Bp = Bp; a.n_cand = B; a.slot0 = 1; a.n_slots = N_NODES; a.n_nodes = N_NODES;
            const long long n_img = (long long)B * N_NODES;
            dec_tail3_kernel<<<(unsigned)(n_img < c->sms ? n_img : c->sms), D3_THREADS, D3_SMEM_BYTES, st>>>(a);
            LAUNCH_CHECK();
        }
    }
    // ---- 7. matched latents + auxiliary heads (base_gcp.py:234-262,361-374)
    float* seq = io->model_enc_seq ? io->model_enc_seq : c->seq;
    {
        const size_t n = (size_t)B * MAX_LEN * (NZ_ENC / 4);
        gather_frames_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e_df, c->frame_node, c->end_ind, B, N_NODES, MAX_LEN,
                                                                           NZ_ENC / 4, seq, c->seqb.p);
        LAUNCH_CHECK();
        float* reg = io->regressed_state ? io->regressed_state : w.reg;
        CHECK(run_pair_heads(c, st, c->end_ind, B, nullptr, reg));
        train_pairs_kernel<<<256, 256, 0, st>>>(w.enc_seq, seq, (const long long*)io->inv_t0, (const long long*)io->inv_t1,
                                                (const long long*)io->cost_start, (const long long*)io->cost_end, B, T, c->pairs.p);
        LAUNCH_CHECK();
        CHECK(mlp_body(c, st, c->inv_mdl, 128, flat, {seg(c->pairs, 0, 256, ROW_LEVEL, 0)}));
        CHECK(gemm(c, st, 128, flat, {seg(c->tb, 0, c->inv_mdl.mid_k)}, c->inv_mdl.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, w.inv_pred, 2, 2)));
        CHECK(mlp_body(c, st, c->cost_mdl, 128, flat, {seg(c->pairs, 0, 256, ROW_LEVEL, 128)}));
        CHECK(gemm(c, st, 128, flat, {seg(c->tb, 0, c->cost_mdl.mid_k)}, c->cost_mdl.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, w.cost_pred, 1, 1)));
        if (io->inv_actions) GCP_CUDA_CHECK(cudaMemcpyAsync(io->inv_actions, w.inv_pred, (size_t)B * 2 * 4, cudaMemcpyDeviceToDevice, st));
        if (io->cost_pred) GCP_CUDA_CHECK(cudaMemcpyAsync(io->cost_pred, w.cost_pred, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
        // ---- 8. reconstruction NLL of every real frame under the node bound to it, KL, scalar losses
        float* nll_bt = io->nll_per_frame ? io->nll_per_frame : w.nll_bt;
#if GCP_VERIFY
        if (c->use_ref) {
            // SIMT path (verification build): decoder tail + NLL of one frame per block
            TailNllArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x3 = w.x3.p; ta.skip_up = c->skip_up; ta.w4p = c->w4p; ta.w5p = c->w5p; ta.b4 = c->b4; ta.b5 = c->b5;
            ta.traj = io->traj_seq; ta.pad_mask = io->pad_mask; ta.end_ind = c->end_ind; ta.frame_node = c->frame_node;
            ta.Bp = Bp; ta.T = T; ta.lcap = MAX_LEN; ta.root_node = (1 << (DEPTH - 1)) - 1; ta.nll_bt = nll_bt;
            dec_tail_nll_kernel<<<B * T, 256, TN_SMEM_BYTES, st>>>(ta);
            LAUNCH_CHECK();
        } else
#endif
        {
            // tcgen05 path: per 64-node chunk the tail kernel runs twice with the raw head (15 mixture-mean logits, 15
            // log-scales; fp32 [seq][node][pixel][16]), then one block per frame evaluates the mixture NLL of the frame
            // under the node bound to it.  All 255 nodes are decoded (half of them are bound to a frame); at 21 ns per
            // node image that is cheaper than gathering the bound ones.
            for (int n0 = 0; n0 < N_NODES; n0 += 64) {
                const int nc = std::min(64, N_NODES - n0);
                for (int half = 0; half < 2; ++half) {
                    DecTail3Args a;
                    memset(&a, 0, sizeof(a));
                    a.x3 = w.x3.p + (size_t)n0 * Bp * 4096; a.s4 = c->s4; a.s4_stride = 256 * 64;
                    a.w4 = c->z4; a.w5 = half ? c->z5s : c->z5m; a.b5h = half ? c->b5s : c->b5m;
                    a.raw = half ? w.raw_ls : w.raw_mu;
                    a.Bp = Bp; a.n_cand = B; a.slot0 = 1 + n0; a.n_slots = nc; a.n_nodes = N_NODES;
                    const long long n_img = (long long)B * nc;
                    dec_tail3_raw_kernel<<<(unsigned)(n_img < c->sms ? n_img : c->sms), D3_THREADS, D3_SMEM_BYTES, st>>>(a);
                    LAUNCH_CHECK();
                }
                RawNllArgs ra;
                memset(&ra, 0, sizeof(ra));
                ra.raw_mu = w.raw_mu; ra.raw_ls = w.raw_ls; ra.traj = io->traj_seq; ra.pad_mask = io->pad_mask;
                ra.end_ind = c->end_ind; ra.frame_node = c->frame_node; ra.T = T; ra.lcap = MAX_LEN;
                ra.root_node = (1 << (DEPTH - 1)) - 1; ra.node0 = n0; ra.n_chunk = nc; ra.nll_bt = nll_bt;
                dlm_nll_raw_kernel<<<B * T, 256, 0, st>>>(ra);
                LAUNCH_CHECK();
            }
        }
        float* kl_b = io->kl_per_seq ? io->kl_per_seq : w.kl_b;
        kl_seq_kernel<<<B, 256, 0, st>>>(q_mu, q_ls, p_mu, p_ls, N_NODES * NZ_VAE, kl_b);
        LAUNCH_CHECK();
        LossArgs la;
        memset(&la, 0, sizeof(la));
        la.B = B; la.T = T; la.n_nodes = N_NODES;
        la.logits = c->logits; la.logits_ld = 256; la.end_ind = c->end_ind;
        la.nll_bt = nll_bt; la.kl_b = kl_b; la.existence = exist; la.keep = w.keep;
        la.reg = reg; la.states = io->states; la.pad_mask = io->pad_mask;
        la.inv_pred = w.inv_pred; la.actions = io->actions; la.inv_t0 = (const long long*)io->inv_t0;
        la.cost_pred = w.cost_pred; la.cost_target = io->cost_target;
        if (!io->cost_target) {
            path_length_kernel<<<B, 256, 0, st>>>(io->traj_seq, (const long long*)io->cost_start, (const long long*)io->cost_end, T,
                                                  w.cost_tgt);
            LAUNCH_CHECK();
            la.cost_target = w.cost_tgt;
        }
        la.frame_elems = (double)T * 3072.0;
        la.losses = io->losses;
        loss_finalize_kernel<<<1, 256, 0, st>>>(la);
        LAUNCH_CHECK();
    }
    return 0;
}

// Sequential GCP rollout (config 3).  Latent rows are time-major: slot t (rows [t*Bp, (t+1)*Bp)) holds x_t, the latent
// of frame t (x_0 = e_0), slot 200 the goal latent.  Per step t = 0..198, all on rows = Bp:
//   prior(x_t) -> (mu, log sigma) -> zeta = exp(log sigma) * z_t + mu            (row-MLP GEMMs, EPI_GN / EPI_REPARAM)
//   embed(cat(x_t, zeta)) + context term                                           (EPI_LINEAR + per-candidate row bias)
//   3 x LSTMCell(1024): gates GEMM [x | h_prev] (K = 2048, N = 4096) + fused cell update, c kept in fp32 in place,
//                       h ping-pongs between two bf16 arrays by step parity       (EPI_LSTM)
//   x_{t+1} = W_o h_2 + b_o -> slot t+1                                             (EPI_LINEAR)
extern "C" int gcpb200_seq_rollout(gcpb200_ctx* c, const gcpb200_seq_io* io, void* stream) {
    if (!io) {
        gcp_set_error("null io");
        return -1;
    }
    CHECK(check_ready(c, io->B));
    if (c->model != GCPB200_MODEL_SEQUENTIAL) {
        gcp_set_error("gcpb200_seq_rollout needs a context created with model = GCPB200_MODEL_SEQUENTIAL");
        return -1;
    }
    if (!io->I_0 || !io->I_g || !io->z) {
        gcp_set_error("gcpb200_seq_rollout: I_0, I_g and z are required");
        return -1;
    }
    if ((io->mu == nullptr) != (io->log_sigma == nullptr)) {
        gcp_set_error("mu and log_sigma must be given together");
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int B = io->B, Bp = (B + 127) / 128 * 128, T = c->max_len - 1, H = c->lstm_hid;
    const LevelGeom flat = {Bp, 0, c->depth};
    const int goal_row0 = c->max_len * Bp;
    const SeqW& S = c->seqw;

    ProfScope total_scope(c, st, 5);
    ProfScope* scope = new ProfScope(c, st, 0);
    {
        CommonIO cio = {io->I_0, io->I_g, io->images_shared, io->end_ind, io->seed, B, io->e_0, io->e_g, io->seq_len_logits,
                        io->end_ind_out};
        CHECK(run_encoder_length(c, st, cio, Bp, goal_row0));
    }
    const std::vector<Seg> ctx_in = {seg(c->lat, 0, NZ_ENC, ROW_LEVEL, 0), seg(c->lat, 0, NZ_ENC, ROW_LEVEL, goal_row0)};
    delete scope;
    scope = new ProfScope(c, st, 1);
    auto recurrence = [&](cudaStream_t st) -> int {
    // context term of the embed layer (constant over the steps) and the initial LSTM state (InitLSTMCell.init_state)
    CHECK(gemm(c, st, Bp, flat, ctx_in, S.embed_ctx, 256, EPI_LINEAR, epi_linear(ACT_NONE, nullptr, 0, c->ctxb, H, H)));
    CHECK(mlp_body(c, st, S.init, Bp, flat, ctx_in));
    {
        EpiParams e = epi_linear(ACT_NONE, c->hs[0].p, 3 * H, c->cs, 3 * H, 6 * H);
        e.split_col = 3 * H;
        CHECK(gemm(c, st, Bp, flat, {seg(c->tb, 0, S.init.mid_k)}, S.init.head, 256, EPI_LINEAR, e));
    }
    for (int t = 0; t < T; ++t) {
        const DevBuf& hcur = c->hs[t & 1];
        const DevBuf& hnxt = c->hs[(t & 1) ^ 1];
        const Seg x_t = seg(c->lat, 0, NZ_ENC, ROW_LEVEL, t * Bp);
        CHECK(mlp_body(c, st, S.prior, Bp, flat, {x_t}));
        {
            EpiParams e;
            memset(&e, 0, sizeof(e));
            e.z = io->z; e.n_cand = B; e.nz = NZ_VAE; e.z_node = t; e.z_nodes = T;
            e.out_bf16 = c->zeta.p; e.out_bf16_ld = NZ_VAE;
            e.mu_out = io->mu; e.ls_out = io->log_sigma;
            CHECK(gemm(c, st, Bp, flat, {seg(c->tb, 0, S.prior.mid_k)}, S.prior.head, 256, EPI_REPARAM, e));
        }
        {
            EpiParams e = epi_linear(ACT_NONE, c->xa.p, H, nullptr, 0, H);
            e.rowbias = c->ctxb;
            e.rowbias_ld = H;
            CHECK(gemm(c, st, Bp, flat, {x_t, seg(c->zeta, 0, NZ_VAE)}, S.embed_main, 256, EPI_LINEAR, e));
        }
        for (int i = 0; i < N_LSTM; ++i) {
            EpiParams e;
            memset(&e, 0, sizeof(e));
            e.c_f32 = c->cs; e.c_f32_ld = 3 * H; e.c_prev_col0 = i * H;
            e.out_bf16 = hnxt.p + i * H; e.out_bf16_ld = 3 * H;
            e.hidden = H;
            const Seg xin = i == 0 ? seg(c->xa, 0, H) : seg(hnxt, (i - 1) * H, H);
            CHECK(gemm(c, st, Bp, flat, {xin, seg(hcur, i * H, H)}, S.lstm[i], 256, EPI_LSTM, e));
        }
        const size_t o = (size_t)(t + 1) * Bp * NZ_ENC;
        CHECK(gemm(c, st, Bp, flat, {seg(hnxt, 2 * H, H)}, S.out, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, c->lat.p + o, NZ_ENC, c->lat_f32 + o, NZ_ENC, NZ_ENC)));
    }
    return 0;
    };
    if (c->seq_graph_on && c->seq_warm && !c->use_ref) {
        gcpb200_ctx::SeqGraph* hit = nullptr;
        for (auto& g : c->seq_graphs)
            if (g.z == io->z && g.mu == io->mu && g.ls == io->log_sigma && g.B == B) hit = &g;
        if (hit == nullptr) {
            const int64_t l0 = c->launches;
            GCP_CUDA_CHECK(cudaStreamBeginCapture(c->seq_stream, cudaStreamCaptureModeRelaxed));
            const int rc = recurrence(c->seq_stream);
            cudaGraph_t graph = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(c->seq_stream, &graph);
            if (rc != 0 || ce != cudaSuccess || graph == nullptr) {
                if (graph) cudaGraphDestroy(graph);
                if (rc == 0) gcp_set_error("sequential rollout: stream capture failed: %s", cudaGetErrorString(ce));
                return -1;
            }
            gcpb200_ctx::SeqGraph g = {io->z, io->mu, io->log_sigma, B, nullptr, c->launches - l0};
            c->launches = l0;
            const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) {
                gcp_set_error("sequential rollout: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
                return -1;
            }
            if (c->seq_graphs.size() >= 8) {
                cudaGraphExecDestroy(c->seq_graphs.front().exec);
                c->seq_graphs.erase(c->seq_graphs.begin());
            }
            c->seq_graphs.push_back(g);
            hit = &c->seq_graphs.back();
        }
        GCP_CUDA_CHECK(cudaEventRecord(c->ev_seq_fork, st));
        GCP_CUDA_CHECK(cudaStreamWaitEvent(c->seq_stream, c->ev_seq_fork, 0));
        GCP_CUDA_CHECK(cudaGraphLaunch(hit->exec, c->seq_stream));
        GCP_CUDA_CHECK(cudaEventRecord(c->ev_seq_join, c->seq_stream));
        GCP_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_seq_join, 0));
        c->launches += hit->launches;
    } else {
        CHECK(recurrence(st));
        c->seq_warm = true;
    }
    delete scope;
    scope = new ProfScope(c, st, 4);
    if (io->encodings) {
        const size_t n = (size_t)B * T * NZ_ENC;
        slot_to_df_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->lat_f32, Bp, B, T, NZ_ENC, NZ_ENC, io->encodings);
        LAUNCH_CHECK();
    }
    delete scope;
    scope = nullptr;
    if (io->images) {
        copy_frame0_kernel<<<(B * 768 + 255) / 256, 256, 0, st>>>(io->I_0, io->images_shared, B, c->max_len, io->images);
        LAUNCH_CHECK();
        CHECK(run_decoder(c, st, io->images_shared, B, Bp, T, io->images + 3072, c->max_len));
    }
    ProfScope asc(c, st, 4);
    if (io->model_enc_seq || io->actions || io->regressed_state) {
        // inputs.end_ind decides how much of cat(e_0, encodings) is kept (phase = 'train' branch, base_gcp.py:238-239)
        const long long* given = reinterpret_cast<const long long*>(io->given_end_ind);
        if (given == nullptr) {
            fill_i64_kernel<<<(B + 255) / 256, 256, 0, st>>>(c->scratch_given, c->max_len - 1, B);
            LAUNCH_CHECK();
            given = c->scratch_given;
        }
        float* seq = io->model_enc_seq ? io->model_enc_seq : c->seq;
        const size_t n = (size_t)B * c->max_len * (NZ_ENC / 4);
        seq_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->lat_f32, given, Bp, B, c->max_len, seq);
        LAUNCH_CHECK();
        if (io->actions || io->regressed_state) {
            const size_t nq = (size_t)B * c->max_len * (NZ_ENC / 4);
            seq_to_b16_rows_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(seq, B, c->max_len, NZ_ENC / 4, c->seqb.p);
            LAUNCH_CHECK();
            CHECK(run_pair_heads(c, st, given, B, io->actions, io->regressed_state));
        }
    }
    return 0;
}

extern "C" int gcpb200_cost_l2_seq(gcpb200_ctx* c, const float* images, int n_frames, const int64_t* end_ind, const float* goal,
                                   int B, int dense, float final_step_weight, float* cost, void* stream) {
    CHECK(check_ready(c, B));
    if (n_frames < 2) {
        gcp_set_error("gcpb200_cost_l2_seq: n_frames = %d", n_frames);
        return -1;
    }
    cost_l2_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(images, nullptr, reinterpret_cast<const long long*>(end_ind),
                                                                            goal, n_frames, n_frames, dense, final_step_weight, cost);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_gather_nodes(gcpb200_ctx* c, const float* src_df, const int32_t* nodes, const int32_t* len, int B,
                                    int row_len, float* dst, void* stream) {
    CHECK(check_ready(c, B));
    if (row_len % 4 || !src_df || !nodes || !len || !dst) {
        gcp_set_error("gcpb200_gather_nodes: bad arguments (row_len must be a multiple of 4)");
        return -1;
    }
    const size_t n = (size_t)B * c->n_nodes * (row_len / 4);
    gather_nodes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src_df, nodes, len, B, c->n_nodes,
                                                                                                      row_len / 4, dst);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_cost_l2_nodes(gcpb200_ctx* c, const float* images_df, const int32_t* nodes, const int32_t* len,
                                     const float* goal, int B, int dense, float final_step_weight, float* cost, void* stream) {
    CHECK(check_ready(c, B));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    len_to_end_kernel<<<(B + 255) / 256, 256, 0, st>>>(len, c->scratch_given, B);
    LAUNCH_CHECK();
    cost_l2_kernel<<<B, 256, 0, st>>>(images_df, nodes, c->scratch_given, goal, c->n_nodes, c->n_nodes, dense, final_step_weight, cost, 0);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_prune_gather(gcpb200_ctx* c, const float* src_df, const int64_t* end_ind, int B, int row_len,
                                    float* dst, void* stream) {
    CHECK(check_ready(c, B));
    if (row_len % 4) {
        gcp_set_error("row_len must be a multiple of 4");
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* ei = reinterpret_cast<const long long*>(end_ind);
    CHECK(compute_frame_map(c, ei, B, st));
    const size_t n = (size_t)B * c->max_len * (row_len / 4);
    gather_frames_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src_df, c->frame_node, ei, B, c->n_nodes, c->max_len, row_len / 4, dst);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_cost_l2(gcpb200_ctx* c, const float* images_df, const int64_t* end_ind, const float* goal, int B,
                               int dense, float final_step_weight, float* cost, void* stream) {
    CHECK(check_ready(c, B));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* ei = reinterpret_cast<const long long*>(end_ind);
    CHECK(compute_frame_map(c, ei, B, st));
    cost_l2_kernel<<<B, 256, 0, st>>>(images_df, c->frame_node, ei, goal, c->n_nodes, c->max_len, dense, final_step_weight, cost);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_cost_learned(gcpb200_ctx* c, const float* e_df, const int64_t* end_ind, int B, const float* goal_seq,
                                    int Lg, float* cost, void* stream) {
    CHECK(check_ready(c, B));
    if (!c->has_cost) {
        gcp_set_error("learned cost needs attach_cost_mdl=1 and cost_mdl.cost_pred.* weights");
        return -1;
    }
    if (Lg < 1 || Lg > c->max_len) {
        gcp_set_error("goal sequence length %d outside [1,%d]", Lg, c->max_len);
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* ei = reinterpret_cast<const long long*>(end_ind);
    const LevelGeom flat = {(B + 127) / 128 * 128, 0, c->depth};
    CHECK(compute_frame_map(c, ei, B, st));
    const size_t n = (size_t)B * c->max_len * (NZ_ENC / 4);
    gather_frames_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e_df, c->frame_node, ei, B, c->n_nodes, c->max_len, NZ_ENC / 4, c->seq);
    LAUNCH_CHECK();
    // pairs inside each candidate's sequence, the pair bridging into the goal sequence, then the goal's own pairs
    const int rows = (B * c->max_len + 127) / 128 * 128;
    make_pairs_kernel<<<(unsigned)(((size_t)rows * 256 + 255) / 256), 256, 0, st>>>(c->seq, ei, goal_seq, B, c->max_len, rows, c->pairs.p);
    LAUNCH_CHECK();
    CHECK(mlp_body(c, st, c->cost_mdl, rows, flat, {seg(c->pairs, 0, 256)}));
    CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->cost_mdl.mid_k)}, c->cost_mdl.head, 128, EPI_LINEAR,
               epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 1, 1)));
    int n_tail = 0;
    if (Lg > 1) {
        // goal-internal pairs: reuse the pair builder on the goal sequence as a 1-candidate batch
        long long* e1 = c->scratch_ei;   // [Lg-1] as "end_ind" of a single pseudo-candidate
        const long long hv = Lg - 1;
        GCP_CUDA_CHECK(cudaMemcpyAsync(e1, &hv, 8, cudaMemcpyHostToDevice, st));
        GCP_CUDA_CHECK(cudaMemsetAsync(c->seq, 0, (size_t)c->max_len * NZ_ENC * 4, st));
        GCP_CUDA_CHECK(cudaMemcpyAsync(c->seq, goal_seq, (size_t)Lg * NZ_ENC * 4, cudaMemcpyDeviceToDevice, st));
        make_pairs_kernel<<<(256 * 256 + 255) / 256, 256, 0, st>>>(c->seq, e1, nullptr, 1, c->max_len, 256, c->pairs.p);
        LAUNCH_CHECK();
        CHECK(mlp_body(c, st, c->cost_mdl, 256, flat, {seg(c->pairs, 0, 256)}));
        CHECK(gemm(c, st, 256, flat, {seg(c->tb, 0, c->cost_mdl.mid_k)}, c->cost_mdl.head, 128, EPI_LINEAR,
                   epi_linear(ACT_NONE, nullptr, 0, c->goal_tail, 1, 1)));
        n_tail = Lg - 1;
    }
    cost_sum_kernel<<<B, 32, 0, st>>>(c->rowcost, ei, c->max_len, c->goal_tail, n_tail, cost);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_cost_pairs(gcpb200_ctx* c, const float* lat, const int32_t* idx1, const int32_t* idx2, int n,
                                  const int32_t* seg_off, int n_seg, float* cost, void* stream) {
    CHECK(check_ready(c, 1));
    if (!c->has_cost) {
        gcp_set_error("learned cost needs attach_cost_mdl=1 and cost_mdl.cost_pred.* weights");
        return -1;
    }
    const int rows = (n + 127) / 128 * 128;
    if (n <= 0 || rows > c->pair_rows) {
        gcp_set_error("gcpb200_cost_pairs: n = %d outside (0, %d]", n, c->pair_rows / 128 * 128);
        return -1;
    }
    if (seg_off != nullptr && n_seg <= 0) {
        gcp_set_error("gcpb200_cost_pairs: seg_off given but n_seg = %d", n_seg);
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const LevelGeom flat = {rows, 0, c->depth};
    make_pairs_idx_kernel<<<(unsigned)(((size_t)rows * 256 + 255) / 256), 256, 0, st>>>(lat, idx1, lat, idx2, n, rows, c->pairs.p);
    LAUNCH_CHECK();
    CHECK(mlp_body(c, st, c->cost_mdl, rows, flat, {seg(c->pairs, 0, 256)}));
    CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->cost_mdl.mid_k)}, c->cost_mdl.head, 128, EPI_LINEAR,
               epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 1, 1)));
    if (seg_off != nullptr) {
        seg_sum_kernel<<<n_seg, 32, 0, st>>>(c->rowcost, seg_off, cost);
        LAUNCH_CHECK();
    } else {
        GCP_CUDA_CHECK(cudaMemcpyAsync(cost, c->rowcost, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

extern "C" int gcpb200_infer_action(gcpb200_ctx* c, const float* img, const float* target_latent, int n, float* action,
                                    float* enc, void* stream) {
    CHECK(check_ready(c, n));
    if (c->model == GCPB200_MODEL_TREE_ADAPTIVE || !c->has_inv) {
        gcp_set_error("gcpb200_infer_action needs inv_mdl.action_pred.* weights (attach_inv_mdl)");
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int rows = (n + 127) / 128 * 128;
    const LevelGeom flat = {rows, 0, c->depth};
    // encoder on the current image(s): latent rows 0..n-1 of the workspace (the skip maps it also writes are scratch
    // that every rollout recomputes)
    encoder_kernel<<<dim3(n, 1), ENC_THREADS, 0, st>>>(img, img, c->enc, c->lat_f32, c->lat.p, 0, 0, c->s0, c->s2, c->s2b.p);
    LAUNCH_CHECK();
    if (enc) GCP_CUDA_CHECK(cudaMemcpyAsync(enc, c->lat_f32, (size_t)n * NZ_ENC * 4, cudaMemcpyDeviceToDevice, st));
    make_pairs_idx_kernel<<<(unsigned)(((size_t)rows * 256 + 255) / 256), 256, 0, st>>>(c->lat_f32, nullptr, target_latent, nullptr, n,
                                                                                        rows, c->pairs.p);
    LAUNCH_CHECK();
    CHECK(mlp_body(c, st, c->inv_mdl, rows, flat, {seg(c->pairs, 0, 256)}));
    CHECK(gemm(c, st, rows, flat, {seg(c->tb, 0, c->inv_mdl.mid_k)}, c->inv_mdl.head, 128, EPI_LINEAR,
               epi_linear(ACT_NONE, nullptr, 0, c->rowcost, 2, 2)));
    GCP_CUDA_CHECK(cudaMemcpyAsync(action, c->rowcost, (size_t)n * 2 * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

extern "C" int gcpb200_topk(gcpb200_ctx* c, const float* cost, int N, int k, int32_t* idx, float* val, void* stream) {
    if (!c || N <= 0 || k <= 0 || k > N) {
        gcp_set_error("gcpb200_topk: bad arguments (N %d, k %d)", N, k);
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (k > c->topk_cap) {
        // rare (first call / larger k): cudaFree synchronises the device, so no in-flight kernel still reads the old array
        if (c->topk_sel) GCP_CUDA_CHECK(cudaFree(c->topk_sel));
        c->topk_sel = nullptr;
        c->topk_cap = 0;
        const int cap = std::max(k, 8192);
        GCP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&c->topk_sel), (size_t)cap * sizeof(unsigned long long)));
        c->topk_cap = cap;
    }
    topk_select_kernel<<<1, TOPK_SELECT_THREADS, 0, st>>>(cost, N, k, c->topk_sel);
    LAUNCH_CHECK();
    topk_rank_kernel<<<(k + 255) / 256, 256, 0, st>>>(c->topk_sel, k, cost, idx, val);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_refit(gcpb200_ctx* c, const float* z, const int32_t* elite_idx, int k, float* mean, float* stdv,
                             void* stream) {
    if (!c || k <= 0) {
        gcp_set_error("gcpb200_refit: bad arguments");
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int per = c->n_nodes * NZ_VAE;
    const int S = std::max(1, std::min(REFIT_SPLITS, k / 8));
    refit_partial_kernel<<<dim3((per / 4 + 127) / 128, S), 128, 0, st>>>(z, elite_idx, k, per, c->refit_part);
    LAUNCH_CHECK();
    refit_final_kernel<<<(per + 255) / 256, 256, 0, st>>>(z, elite_idx, k, per, S, c->refit_part, mean, stdv);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_sample_noise(gcpb200_ctx* c, const float* mean, const float* stdv, float std_scalar, uint64_t seed,
                                    uint64_t first_candidate_id, int B, float clip, float* z, void* stream) {
    if (!c || B <= 0) {
        gcp_set_error("gcpb200_sample_noise: bad arguments");
        return -1;
    }
    const int per = c->n_nodes * NZ_VAE;
    sample_noise_kernel<<<dim3(B, (per / 4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mean, stdv, std_scalar, seed, first_candidate_id, nullptr, B, per, clip, z);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_sample_noise_ids(gcpb200_ctx* c, const float* mean, const float* stdv, float std_scalar,
                                        uint64_t seed, const int32_t* ids, int B, float clip, float* z, void* stream) {
    if (!c || B <= 0 || !ids) {
        gcp_set_error("gcpb200_sample_noise_ids: bad arguments");
        return -1;
    }
    const int per = c->n_nodes * NZ_VAE;
    sample_noise_kernel<<<dim3(B, (per / 4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        mean, stdv, std_scalar, seed, 0ULL, ids, B, per, clip, z);
    LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// DTW family: no weights involved, so these only need a context (device + launch accounting)
// ---------------------------------------------------------------------------------------------
extern "C" int gcpb200_cdist_mean(gcpb200_ctx* c, const float* x, const float* y, int B, int n, int m, int dim, float* out,
                                  void* stream) {
    if (!c || !x || !y || !out || B <= 0 || n <= 0 || m <= 0 || dim <= 0 || B > 65535) {
        gcp_set_error("gcpb200_cdist_mean: bad arguments (B=%d n=%d m=%d dim=%d)", B, n, m, dim);
        return -1;
    }
    if (dim % 4 != 0 || (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) % 16 != 0) {
        gcp_set_error("gcpb200_cdist_mean: vectors must be 16-byte aligned with a length that is a multiple of 4 (dim=%d)", dim);
        return -1;
    }
    dim3 grid((m + CD_T - 1) / CD_T, (n + CD_T - 1) / CD_T, B);
    cdist_mean_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, m, dim, out);
    LAUNCH_CHECK();
    return 0;
}

extern "C" size_t gcpb200_soft_dtw_workspace(int B, int r, int c) {
    return (B <= 0 || r <= 0 || c <= 0) ? 0 : (size_t)2 * B * r * c * sizeof(double);
}

extern "C" int gcpb200_soft_dtw(gcpb200_ctx* c, const float* cost, float temp, const int64_t* end_inds, int B, int r, int cols,
                                void* workspace, float* w, float* w_bf, float* rowsum_max, void* stream) {
    if (!c || !cost || !workspace || !w || B <= 0 || r <= 0 || cols <= 0 || B > 32767) {
        gcp_set_error("gcpb200_soft_dtw: bad arguments (B=%d r=%d c=%d)", B, r, cols);
        return -1;
    }
    if (r < cols) {      // probabilistic_dtw.py:35 `assert r >= c`
        gcp_set_error("gcpb200_soft_dtw: needs at least as many nodes as frames (r=%d < c=%d)", r, cols);
        return -1;
    }
    if (cols > 3072) {
        gcp_set_error("gcpb200_soft_dtw: c=%d > 3072 frames", cols);
        return -1;
    }
    if (!(temp != 0.f)) {
        gcp_set_error("gcpb200_soft_dtw: temperature must be non-zero");
        return -1;
    }
    int depth = 0;
    if (w_bf) {
        while ((1 << depth) - 1 < r) ++depth;
        if ((1 << depth) - 1 != r) {
            gcp_set_error("gcpb200_soft_dtw: breadth-first output needs r = 2^d - 1 tree nodes (r=%d)", r);
            return -1;
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* ei = reinterpret_cast<const long long*>(end_inds);
    double* accum = reinterpret_cast<double*>(workspace);
    const int threads = std::min(1024, (cols + 31) / 32 * 32);
    soft_dtw_sweep_kernel<<<2 * B, threads, 2 * cols * sizeof(double), st>>>(cost, temp, ei, B, r, cols, accum);
    LAUNCH_CHECK();
    const size_t n = (size_t)B * r * cols;
    soft_dtw_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cost, temp, ei, accum, B, r, cols, w);
    LAUNCH_CHECK();
    if (rowsum_max) {
        GCP_CUDA_CHECK(cudaMemsetAsync(rowsum_max, 0, sizeof(float), st));
        soft_dtw_rowsum_max_kernel<<<(B * r + 7) / 8, 256, 0, st>>>(w, B * r, cols, rowsum_max);
        LAUNCH_CHECK();
    }
    if (w_bf) {
        binding_normalize_kernel<<<dim3((cols + 127) / 128, B), 128, 0, st>>>(w, B, r, cols, depth, 1e-7f, w_bf);
        LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int gcpb200_dtw(gcpb200_ctx* c, const void* cost, int cost_is_f64, const int64_t* end_ind, int B, int r, int cols,
                           double* acc, double* dist, int32_t* path_p, int32_t* path_q, int32_t* path_len, int32_t* match_inds,
                           void* stream) {
    if (!c || !cost || !acc || !dist || !path_p || !path_q || !path_len || B <= 0 || r <= 0 || cols <= 0) {
        gcp_set_error("gcpb200_dtw: bad arguments (B=%d r=%d c=%d)", B, r, cols);
        return -1;
    }
    if (r > 2048) {
        gcp_set_error("gcpb200_dtw: r=%d > 2048 rows", r);
        return -1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* ei = reinterpret_cast<const long long*>(end_ind);
    const int threads = std::min(1024, (std::min(r, cols) + 31) / 32 * 32);
    const size_t smem = (size_t)3 * r * sizeof(double);
    if (cost_is_f64)
        dtw_wavefront_kernel<double><<<B, threads, smem, st>>>(reinterpret_cast<const double*>(cost), ei, r, cols, acc, dist,
                                                                path_p, path_q, path_len, match_inds);
    else
        dtw_wavefront_kernel<float><<<B, threads, smem, st>>>(reinterpret_cast<const float*>(cost), ei, r, cols, acc, dist,
                                                               path_p, path_q, path_len, match_inds);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_gather_rows(gcpb200_ctx* c, const float* src, const int32_t* idx, int n_out, int row_floats, float* out,
                                   void* stream) {
    if (!c || !src || !idx || !out || n_out <= 0 || row_floats <= 0 || row_floats % 4 != 0) {
        gcp_set_error("gcpb200_gather_rows: bad arguments (rows must be a multiple of 4 floats)");
        return -1;
    }
    const size_t n = (size_t)n_out * (row_floats / 4);
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(src), idx, n_out, row_floats / 4, reinterpret_cast<float4*>(out));
    LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Optimiser step (RAdam / Adam + gradient-norm clipping)
// ---------------------------------------------------------------------------------------------
extern "C" int gcpb200_sq_norm(gcpb200_ctx* c, const float* x, int64_t n, double* acc, void* stream) {
    if (!c || !x || !acc || n < 0) {
        gcp_set_error("gcpb200_sq_norm: bad arguments");
        return -1;
    }
    if (n == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long blocks = (n + 256 * 8 - 1) / (256 * 8);
    sq_norm_kernel<<<(unsigned)(blocks < 4LL * c->sms ? blocks : 4LL * c->sms), 256, 0, st>>>(x, (long long)n, acc);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gcpb200_optim_step(gcpb200_ctx* c, int kind, float* p, const float* g, float* m, float* v, int64_t n, double lr,
                                  double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                  const double* grad_sq_norm, float max_norm, void* stream) {
    if (!c || !p || !g || !m || !v || n < 0 || step < 1 || (kind != GCPB200_OPT_ADAM && kind != GCPB200_OPT_RADAM)) {
        gcp_set_error("gcpb200_optim_step: bad arguments (kind %d, n %lld, step %lld)", kind, (long long)n, (long long)step);
        return -1;
    }
    if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) {
        gcp_set_error("gcpb200_optim_step: p, g, m, v must be 16-byte aligned");
        return -1;
    }
    if (n == 0) return 0;
    OptimArgs a;
    memset(&a, 0, sizeof(a));
    a.p = p; a.g = g; a.m = m; a.v = v; a.n = (long long)n;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.grad_sq_norm = grad_sq_norm; a.max_norm = max_norm;
    // per-step scalars in double, as the Python floats of the reference
    const double b1 = beta1, b2 = beta2, t = (double)step;
    if (kind == GCPB200_OPT_RADAM) {
        // blox/torch/radam.py:56-68
        const double beta2_t = pow(b2, t);
        const double n_sma_max = 2.0 / (1.0 - b2) - 1.0;
        const double n_sma = n_sma_max - 2.0 * t * beta2_t / (1.0 - beta2_t);
        double step_size;
        if (n_sma >= 5.0)
            step_size = sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_sma_max - 4.0) * (n_sma - 2.0) / n_sma * n_sma_max /
                             (n_sma_max - 2.0)) / (1.0 - pow(b1, t));
        else
            step_size = 1.0 / (1.0 - pow(b1, t));
        a.radam = 1;
        a.rectified = n_sma >= 5.0;
        a.step_size = (float)(step_size * lr);
        a.decay = (float)(weight_decay * lr);
    } else {
        a.step_size = (float)(lr / (1.0 - pow(b1, t)));
        a.sqrt_bc2 = (float)sqrt(1.0 - pow(b2, t));
        a.decay = (float)weight_decay;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long blocks = (n / 4 + 255) / 256 + 1;
    optim_step_kernel<<<(unsigned)(blocks < 8LL * c->sms ? blocks : 8LL * c->sms), 256, 0, st>>>(a);
    LAUNCH_CHECK();
    return 0;
}

