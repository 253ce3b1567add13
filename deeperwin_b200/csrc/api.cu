// C ABI of libdpe_b200.so (include/dpe_b200.h): handle, parameter layout, workspace planning and the
// kernel sequences of log psi^2, E_loc and the Metropolis step.
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>
#include "dpe_internal.cuh"

namespace dpe {

static thread_local char g_err[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return DPE_OK;
    return set_error(DPE_ERR_CUDA, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- dims helpers ---------------------------------------------------------------------------------
static int d_one_in(const dpe_dims &d, int it) { return it == 0 ? 4 * d.n_ion : d.n_hidden_one_el[it - 1]; }
static int d_pair_in(const dpe_dims &d, int it) { return it == 0 ? 1 : d.n_hidden_two_el[it - 1]; }
static int d_eion_in(const dpe_dims &d, int it) { return it == 0 ? 4 : d.n_hidden_two_el[it - 1]; }

static int validate_dims(const dpe_dims &d) {
    if (d.n_el < 2 || d.n_el > 64) return set_error(DPE_ERR_UNSUPPORTED, "n_el=%d outside [2, 64]", d.n_el);
    if (d.n_up < 1 || d.n_up >= d.n_el) return set_error(DPE_ERR_UNSUPPORTED, "need at least one electron of each spin (n_up=%d)", d.n_up);
    if (d.n_ion < 1 || d.n_ion > 256) return set_error(DPE_ERR_UNSUPPORTED, "n_ion=%d outside [1, 256]", d.n_ion);
    if (d.n_iterations < 1 || d.n_iterations > DPE_MAX_ITER) return set_error(DPE_ERR_UNSUPPORTED, "n_iterations=%d", d.n_iterations);
    if (d.emb_dim < 8 || d.emb_dim > 32 || (d.emb_dim & 7)) return set_error(DPE_ERR_UNSUPPORTED, "emb_dim=%d must be a multiple of 8 in [8, 32]", d.emb_dim);
    if (d.n_ion_features < 1 || d.n_dets < 1) return set_error(DPE_ERR_UNSUPPORTED, "n_ion_features / n_dets must be positive");
    if (d.z_max < d.z_min) return set_error(DPE_ERR_ARG, "z_max < z_min");
    if (d.use_taos != 0 && d.use_taos != 1) return set_error(DPE_ERR_ARG, "use_taos must be 0 or 1");
    for (int it = 0; it < d.n_iterations; ++it) {
        if (d.n_hidden_one_el[it] < 4 || (d.n_hidden_one_el[it] & 3)) return set_error(DPE_ERR_UNSUPPORTED, "n_hidden_one_el[%d]=%d must be a multiple of 4", it, d.n_hidden_one_el[it]);
        if (it + 1 < d.n_iterations && (d.n_hidden_two_el[it] < 4 || d.n_hidden_two_el[it] > 32 || (d.n_hidden_two_el[it] & 3)))
            return set_error(DPE_ERR_UNSUPPORTED, "n_hidden_two_el[%d]=%d must be a multiple of 4 in [4, 32]", it, d.n_hidden_two_el[it]);
        int din = d_one_in(d, it);
        if (3 * din + d.emb_dim + d_eion_in(d, it) == d.n_hidden_one_el[it])
            return set_error(DPE_ERR_UNSUPPORTED, "h_el_%d would take the residual branch (mlp.py:13-16); not implemented", it);
    }
    return DPE_OK;
}

static void add_leaf(dpe_model *m, int rows, int cols) {
    Leaf &l = m->leaves[m->n_leaves++];
    l.off = m->n_params; l.rows = rows; l.cols = cols; l.size = (int64_t)rows * cols;
    m->n_params += l.size;
}

static void build_leaves(dpe_model *m) {
    const dpe_dims &d = m->dims;
    m->n_leaves = 0; m->n_params = 0;
    add_leaf(m, d.z_max - d.z_min + 1, d.n_ion_features);
    for (int it = 0; it < d.n_iterations; ++it) {
        int din = d_one_in(d, it), dP = d_pair_in(d, it), dE = d_eion_in(d, it), dout = d.n_hidden_one_el[it];
        add_leaf(m, dP, d.emb_dim); add_leaf(m, 1, d.emb_dim);      // w_same
        add_leaf(m, dP, d.emb_dim); add_leaf(m, 1, d.emb_dim);      // w_diff
        add_leaf(m, din, d.emb_dim); add_leaf(m, 1, d.emb_dim);     // h_map
        add_leaf(m, d.n_ion_features, dE); add_leaf(m, 1, dE);      // h_ion_map
        add_leaf(m, 3 * din + d.emb_dim + dE, dout); add_leaf(m, 1, dout);   // h_el
        if (it + 1 < d.n_iterations) {
            int d2 = d.n_hidden_two_el[it];
            add_leaf(m, dP, d2); add_leaf(m, 1, d2);                // h_same
            add_leaf(m, dP, d2); add_leaf(m, 1, d2);                // h_diff
            add_leaf(m, dE, d2); add_leaf(m, 1, d2);                // h_el_ion
        }
    }
    if (d.use_taos) return;       // TAO heads: backflows / exponents come from the geometry cache (dpe_model_set_tao_cache)
    int cols = d.n_dets * d.n_el, dl = d.n_hidden_one_el[d.n_iterations - 1];
    add_leaf(m, dl, cols); add_leaf(m, dl, cols);                   // bf_up, bf_dn
    for (int q = 0; q < 4; ++q) add_leaf(m, d.n_ion, cols);         // alpha_up, alpha_dn, weights_up, weights_dn
}

static void bind_views(dpe_model *m) {
    const dpe_dims &d = m->dims;
    int li = 0;
    auto next = [&]() { return m->params + m->leaves[li++].off; };
    auto dense = [&](Dense &L, int din, int dout) { L.w = next(); L.b = next(); L.din = din; L.dout = dout; };
    m->h_ion_emb = next();
    float *dv = m->derived;
    for (int it = 0; it < d.n_iterations; ++it) {
        IterParams &p = m->it[it];
        p.d_in = d_one_in(d, it); p.dP = d_pair_in(d, it); p.dE = d_eion_in(d, it); p.d_out = d.n_hidden_one_el[it];
        p.k_main = p.d_in + d.emb_dim + p.dE;
        dense(p.w_same, p.dP, d.emb_dim);
        dense(p.w_diff, p.dP, d.emb_dim);
        dense(p.h_map, p.d_in, d.emb_dim);
        dense(p.h_ion_map, d.n_ion_features, p.dE);
        dense(p.h_el, 3 * p.d_in + d.emb_dim + p.dE, p.d_out);
        if (it + 1 < d.n_iterations) {
            int d2 = d.n_hidden_two_el[it];
            dense(p.h_same, p.dP, d2);
            dense(p.h_diff, p.dP, d2);
            dense(p.h_el_ion, p.dE, d2);
            p.pair_next = d2; p.eion_next = d2;
        }
        p.w_main = dv; dv += (size_t)p.k_main * p.d_out;
        p.w_mean = dv; dv += (size_t)2 * p.d_in * p.d_out;
        p.him = dv; dv += (size_t)d.n_ion * p.dE;
        dv = m->derived + align_up((dv - m->derived) * sizeof(float)) / sizeof(float);
    }
    if (d.use_taos) return;
    m->bf_w[0] = next(); m->bf_w[1] = next();
    m->alpha[0] = next(); m->alpha[1] = next();
    m->env_w[0] = next(); m->env_w[1] = next();
    size_t nsp = (size_t)d.n_ion * d.n_dets * d.n_el;
    m->sp_alpha[0] = dv; dv += nsp;
    m->sp_alpha[1] = dv; dv += nsp;
}

static size_t derived_floats(const dpe_dims &d) {
    size_t n = 0;
    for (int it = 0; it < d.n_iterations; ++it) {
        int din = d_one_in(d, it), dE = d_eion_in(d, it), dout = d.n_hidden_one_el[it];
        n += (size_t)(din + d.emb_dim + dE) * dout + (size_t)2 * din * dout + (size_t)d.n_ion * dE;
        n = align_up(n * sizeof(float)) / sizeof(float);
    }
    n += 2 * (size_t)d.n_ion * d.n_dets * d.n_el;
    return n + 64;
}

// ---- workspace ------------------------------------------------------------------------------------
static void plan(const dpe_dims &d, int Bc, int C, WsLayout &L) {
    const int N = d.n_el, CP = C > 1 ? 3 : 1, CE = C > 1 ? 5 : 1;
    int ldx = 0, max_din = 0, max_dout = 0;
    for (int it = 0; it < d.n_iterations; ++it) {
        int din = d_one_in(d, it), km = din + d.emb_dim + d_eion_in(d, it);
        if (km > ldx) ldx = km;
        if (din > max_din) max_din = din;
        if (d.n_hidden_one_el[it] > max_dout) max_dout = d.n_hidden_one_el[it];
    }
    if (max_dout > ldx) ldx = max_dout;
    L.ldx = ldx;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off = align_up(off + floats * sizeof(float)); return o; };
    const size_t rows = (size_t)Bc * N * C;
    L.x[0] = take(rows * ldx);
    L.x[1] = take(rows * ldx);
    L.hm = take(rows * d.emb_dim);
    L.mean = take((size_t)Bc * C * 2 * max_din);
    L.add = take((size_t)Bc * C * max_dout);
    L.pw = off;
    for (int it = 0; it < d.n_iterations; ++it) L.pw_it[it] = take((size_t)Bc * N * N * CP * d.emb_dim);
    L.ei = off;
    for (int it = 0; it < d.n_iterations; ++it) L.ei_it[it] = take((size_t)Bc * N * CE * d_eion_in(d, it));
    L.mo = take(rows * d.n_dets * N);
    L.det = take((size_t)Bc * d.n_dets * (C > 1 ? 3 * N + 3 : 2));
    {   // tf32-split Ainv^T tiles for the tensor-core determinant stage (Laplacian mode)
        const size_t NP = det_tc_pad(N, d.n_dets);
        L.ainv = take(C > 1 && NP <= 64 ? (size_t)Bc * d.n_dets * NP * NP * 2 : 0);
    }
    L.tao_g = take(d.use_taos ? rows * d.n_ion * d.n_dets * N : 0);      // per-ion orbital pre-factors h_i . b[ion, orb, det]
    L.epot = take(Bc);
    L.lp = take(Bc);
    L.total_chunk = off;
}

static void plan_mcmc(const dpe_dims &d, int B, WsLayout &L) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    L.r_prop = take((size_t)B * d.n_el * 3 * sizeof(float));
    L.lp_prop = take((size_t)B * sizeof(float));
    L.thr = take((size_t)B * sizeof(float));
    L.new_keys = take((size_t)B * 2 * sizeof(uint32_t));
    L.log_q = take((size_t)B * sizeof(float));
    L.total_mcmc = off;
}

static int max_chunk(const dpe_dims &d, int C, size_t avail, int B) {
    WsLayout L;
    plan(d, B, C, L);
    if (L.total_chunk <= avail) return B;
    int lo = 0, hi = B;       // invariant: plan(lo) fits (lo = 0 trivially), plan(hi) does not
    while (hi - lo > 1) {
        int mid = lo + (hi - lo) / 2;
        plan(d, mid, C, L);
        if (L.total_chunk <= avail) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- kernel sequences -------------------------------------------------------------------------------
// *fused is set when the kernel that ran applied the requested fused epilogue
static int gemm_dispatch(dpe_model *m, const GemmArgs &g, cudaStream_t s, bool *fused) {
    if (fused) *fused = false;
    if (m->gemm_path == 1) {
        int e = launch_gemm_tc(m, g, s);
        if (e != DPE_ERR_UNSUPPORTED) { if (fused) *fused = g.epi != 0; return e; }
        if (g.epi) {                 // shape without a fused epilogue (e.g. the narrow rows kernel): plain tensor-core GEMM
            GemmArgs g0 = g;
            g0.epi = 0;
            e = launch_gemm_tc(m, g0, s);
            if (e != DPE_ERR_UNSUPPORTED) return e;
        }
    }
    return launch_gemm_simt(m, g, s);
}

// site: 0 h_map, 1 spin-mean term, 2 main layer, 3 backflow / TAO projection.  DPE_TC_SIMT_MASK (debug) sends the sites whose bit is
// set through the FP32 SIMT GEMM while the rest stays on the tensor cores (accuracy attribution, tools/parity_table.py).
static int gemm(dpe_model *m, const GemmArgs &g, cudaStream_t s, bool *fused = nullptr, int site = -1) {
    static const int simt_mask = getenv("DPE_TC_SIMT_MASK") ? atoi(getenv("DPE_TC_SIMT_MASK")) : 0;
    if (site >= 0 && (simt_mask >> site & 1)) {
        if (fused) *fused = false;
        return launch_gemm_simt(m, g, s);
    }
    if (!m->profile) return gemm_dispatch(m, g, s, fused);
    dpe_model::ProfRec rec;
    DPE_CUDA(cudaEventCreate(&rec.e0));
    DPE_CUDA(cudaEventCreate(&rec.e1));
    DPE_CUDA(cudaEventRecord(rec.e0, s));
    int e = gemm_dispatch(m, g, s, fused);
    DPE_CUDA(cudaEventRecord(rec.e1, s));
    rec.klass = m->last_gemm_class;
    rec.stage = -1;
    rec.flops = 2.0 * (double)g.M * (double)g.N * (double)g.K;
    m->prof->push_back(rec);
    return e;
}

static GemmArgs plain_gemm(const float *A, int lda, const float *W, int ldw, float *Cc, int ldc, int M, int N, int K) {
    GemmArgs g;
    g.A = A; g.lda = lda; g.a_seg_len = M > 0 ? M : 1; g.a_seg_stride = 0; g.a_seg_off = 0;
    g.W = W; g.ldw = ldw;
    g.C = Cc; g.ldc = ldc; g.c_seg_len = M > 0 ? M : 1; g.c_seg_stride = 0; g.c_seg_off = 0; g.c_col_off = 0;
    g.M = M; g.N = N; g.K = K;
    return g;
}

// Dense layers for the other translation units (grad.cu): C = A W through the model's GEMM path (tensor cores where the weight is registered)
int dense_gemm(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, cudaStream_t s) {
    return gemm(m, plain_gemm(A, lda, W, N, Cc, ldc, M, N, K), s);
}
// C = A W^T with the layer's own weight W [N, K] (registered with tr): the backward data products of the gradient pass
int dense_gemm_t(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, int seg_len, int seg_stride, int seg_off,
                 cudaStream_t s) {
    static const bool off = getenv("DPE_GEMM_NT_TC") && atoi(getenv("DPE_GEMM_NT_TC")) == 0;
    if (off || m->gemm_path != 1) return DPE_ERR_UNSUPPORTED;
    GemmArgs g = plain_gemm(A, lda, W, K, Cc, ldc, M, N, K);
    g.w_tr = 1;
    if (seg_len > 0) {
        g.a_seg_len = g.c_seg_len = seg_len;
        g.a_seg_stride = g.c_seg_stride = seg_stride;
        g.a_seg_off = g.c_seg_off = seg_off;
    }
    return launch_gemm_tc(m, g, s);
}
// ... on the row segment [seg_off, seg_off + seg_len) of every block of seg_stride rows (the spin block of a walker)
int dense_gemm_seg(dpe_model *m, const float *A, int lda, const float *W, float *Cc, int ldc, int M, int N, int K, int seg_len, int seg_stride, int seg_off,
                   cudaStream_t s) {
    GemmArgs g = plain_gemm(A, lda, W, N, Cc, ldc, M, N, K);
    g.a_seg_len = g.c_seg_len = seg_len;
    g.a_seg_stride = g.c_seg_stride = seg_stride;
    g.a_seg_off = g.c_seg_off = seg_off;
    return gemm(m, g, s);
}

// One chunk of Bc walkers through the whole network. C = 1 (forward) or 3N+2 (forward Laplacian).
static int run_chunk(dpe_model *m, const float *r, int Bc, int C, char *ws, const WsLayout &L, float *phase, float *logpsi2,
                     float *grad, float *ekin, float *eloc, float *epot_out, cudaStream_t s) {
    const dpe_dims &d = m->dims;
    const int N = d.n_el, U = d.n_up, D = N - U;
    const int CP = C > 1 ? 3 : 1, CE = C > 1 ? 5 : 1;
    (void)CP;
    float *x[2] = {(float *)(ws + L.x[0]), (float *)(ws + L.x[1])};
    float *hm = (float *)(ws + L.hm), *mean = (float *)(ws + L.mean), *add = (float *)(ws + L.add);
    float *mo = (float *)(ws + L.mo), *det = (float *)(ws + L.det), *epot = (float *)(ws + L.epot);
    size_t pw_off[DPE_MAX_ITER], ei_off[DPE_MAX_ITER];
    for (int it = 0; it < d.n_iterations; ++it) {
        pw_off[it] = (L.pw_it[it] - L.pw) / sizeof(float);
        ei_off[it] = (L.ei_it[it] - L.ei) / sizeof(float);
    }
    float *pw = (float *)(ws + L.pw), *ei = (float *)(ws + L.ei);
    const int ldx = L.ldx;
    const int rows = Bc * N * C;
    int e;
    { StageTimer t(m, ST_FEATURES, s); if ((e = launch_features(m, r, Bc, C, x[0], ldx, C > 1 ? epot : nullptr, s))) return e; }
    { StageTimer t(m, ST_EION, s); if ((e = launch_eion_stream(m, r, Bc, CE, ei, ei_off, s))) return e; }
    { StageTimer t(m, ST_PAIR, s); if ((e = launch_pair_stream(m, r, Bc, C > 1 ? 3 : 1, pw, pw_off, s))) return e; }
    int cur = 0;
    bool mean_ready = false;          // the spin means of x[cur] were already formed by the previous iteration's activation kernel (forward pass)
    for (int it = 0; it < d.n_iterations; ++it) {
        const IterParams &p = m->it[it];
        // h_map: [rows, d_in] x [d_in, emb] -> hm, tanh rule
        {
            StageTimer t(m, ST_HMAP, s);
            GemmArgs gh = plain_gemm(x[cur], ldx, p.h_map.w, d.emb_dim, hm, d.emb_dim, rows, d.emb_dim, p.d_in);
            bool fused_h = false;
            if (C == 1) { gh.epi = 1; gh.n_ch = 1; gh.bias = p.h_map.b; gh.add = nullptr; gh.groups_per_add = 1; }    // forward pass: tanh in the rows-GEMM epilogue
            if ((e = gemm(m, gh, s, &fused_h, 0))) return e;
            if (!fused_h && (e = launch_act(m, hm, d.emb_dim, Bc * N, C, d.emb_dim, p.h_map.b, nullptr, 1, s))) return e;
        }
        // SchNet convolutions fill columns [d_in, k_main)
        { StageTimer t(m, ST_CONV, s); if ((e = launch_conv(m, it, r, Bc, C, hm, pw + pw_off[it], ei + ei_off[it], x[cur], ldx, s))) return e; }
        // spin means and their contribution (shared by all electrons of a walker)
        if (!mean_ready) { StageTimer t(m, ST_MEAN, s); if ((e = launch_mean(m, x[cur], ldx, Bc, C, p.d_in, mean, s))) return e; }
        mean_ready = false;
        {
            StageTimer t(m, ST_MEAN_GEMM, s);
            if ((e = gemm(m, plain_gemm(mean, 2 * p.d_in, p.w_mean, p.d_out, add, p.d_out, Bc * C, p.d_out, 2 * p.d_in), s, nullptr, 1))) return e;
        }
        // main layer
        {
            StageTimer t(m, ST_MAIN, s);
            // bias + spin-mean addend + tanh rule are applied in the epilogue of the CTA-pair tensor-core kernel, where the double
            // buffered accumulators let it run under the next tile's MMAs (gemm_tc.cu); every other kernel leaves `fused`
            // false and k_act does it as a separate HBM-bound pass.
            GemmArgs g = plain_gemm(x[cur], ldx, p.w_main, p.d_out, x[cur ^ 1], ldx, rows, p.d_out, p.k_main);
            static const bool fuse_act = getenv("DPE_FUSE_ACT") != nullptr || tc_pair_mode() > 1;
            if (fuse_act) { g.epi = 1; g.n_ch = C; g.bias = p.h_el.b; g.add = add; g.groups_per_add = N; }
            bool fused = false;
            if ((e = gemm(m, g, s, &fused, 2))) return e;
            if (!fused) {
                // forward pass: the activation kernel also leaves the spin means of its output for the next iteration (no second pass over it)
                static const bool no_act_mean = getenv("DPE_NO_ACT_MEAN") != nullptr;
                if (C == 1 && !no_act_mean) {
                    const bool next = it + 1 < d.n_iterations && m->it[it + 1].d_in == p.d_out;
                    if ((e = launch_act_mean_fwd(m, x[cur ^ 1], ldx, Bc, p.d_out, p.h_el.b, add, next ? mean : nullptr, s))) return e;
                    mean_ready = next;
                } else if ((e = launch_act(m, x[cur ^ 1], ldx, Bc * N, C, p.d_out, p.h_el.b, add, N, s))) return e;
            }
        }
        cur ^= 1;
    }
    const int cols = d.n_dets * N, dl = d.n_hidden_one_el[d.n_iterations - 1];
    {
    StageTimer t_orb(m, ST_ORBITALS, s);
    if (d.use_taos) {
        // transferable atomic orbitals (transferable_atomic_orbitals.py:287-349): one GEMM against the cached backflow matrix
        // for all electrons (both spin types use slice 0, :255-260), then the exponential envelopes and the sum over ions
        float *tg = (float *)(ws + L.tao_g);
        const int gc = d.n_ion * cols;
        if ((e = gemm(m, plain_gemm(x[cur], ldx, m->tao_w, gc, tg, gc, rows, gc, dl), s, nullptr, 3))) return e;
        if ((e = launch_tao_orbitals(m, r, Bc, C, tg, mo, s))) return e;
    } else {
    // backflow factors (envelope_orbitals.py:46-75): spin-up / spin-down electrons use different matrices
    // The envelope is applied in the epilogue of the CTA-pair kernel (double-buffered accumulators: it overlaps the next tile's
    // MMAs); in the single-CTA kernel the same fusion is exposed and slower (measured N2 45.3 vs 41.5 ms/step), there, on the
    // SIMT path and in the forward pass (short packed segments) k_envelope runs as a separate pass.
    static const bool fuse_env_on = getenv("DPE_FUSE_ENVELOPE") != nullptr || (tc_pair_mode() > 1 && !getenv("DPE_NO_FUSE_ENVELOPE"));
    bool want_fused = fuse_env_on && C > 1, env_fused = false;
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool all_fused = true;
        for (int sp = 0; sp < 2; ++sp) {
            GemmArgs g = plain_gemm(x[cur], ldx, m->bf_w[sp], cols, mo, cols, Bc * (sp ? D : U) * C, cols, dl);
            g.a_seg_len = g.c_seg_len = (sp ? D : U) * C;
            g.a_seg_stride = g.c_seg_stride = N * C;
            g.a_seg_off = g.c_seg_off = sp ? U * C : 0;
            if (want_fused) {
                g.epi = 2; g.n_ch = C; g.r = r; g.R = m->R_dev; g.spa = m->sp_alpha[sp]; g.envw = m->env_w[sp];
                g.n_el = N; g.n_ion = d.n_ion; g.el_base = sp ? U : 0;
            }
            bool fused = false;
            if ((e = gemm(m, g, s, &fused, 3))) return e;
            all_fused = all_fused && fused;
            if (want_fused && !fused) break;        // this shape has no fused kernel: redo both spin blocks plainly
        }
        env_fused = want_fused && all_fused;
        if (env_fused || !want_fused) break;
        want_fused = false;
    }
    if (!env_fused && (e = launch_envelope(m, r, Bc, C, mo, s))) return e;
    }
    }
    if ((e = launch_det(m, Bc, C, mo, det, C > 1 ? (float *)(ws + L.ainv) : nullptr, s))) return e;
    { StageTimer t(m, ST_COMBINE, s); if ((e = launch_combine(m, Bc, C, det, epot, phase, logpsi2, grad, ekin, eloc, epot_out, s))) return e; }
    return DPE_OK;
}

static int run_batched(dpe_model *m, const float *r, int B, int C, char *ws, size_t ws_bytes, float *phase, float *logpsi2,
                       float *grad, float *ekin, float *eloc, float *epot_out, cudaStream_t s) {
    if (!m->params_set || !m->geom_set) return set_error(DPE_ERR_STATE, "set_params and set_geometry must be called first");
    if (m->dims.use_taos && !m->tao_set) return set_error(DPE_ERR_STATE, "this model uses transferable atomic orbitals: set_tao_cache must be called first");
    const dpe_dims &d = m->dims;
    int chunk = max_chunk(d, C, ws_bytes, B);
    if (chunk < 1) return set_error(DPE_ERR_WORKSPACE, "workspace of %zu bytes cannot hold one walker", ws_bytes);
    const int K = 3 * d.n_el;
    for (int off = 0; off < B; off += chunk) {
        int Bc = B - off < chunk ? B - off : chunk;
        WsLayout L;
        plan(d, Bc, C, L);
        float *lp = logpsi2 ? logpsi2 + off : (float *)(ws + L.lp);
        int e = run_chunk(m, r + (size_t)off * d.n_el * 3, Bc, C, ws, L, phase ? phase + off : nullptr, lp,
                          grad ? grad + (size_t)off * K : nullptr, ekin ? ekin + off : nullptr, eloc ? eloc + off : nullptr,
                          epot_out ? epot_out + off : nullptr, s);
        if (e) return e;
    }
    return DPE_OK;
}

}  // namespace dpe

using namespace dpe;

extern "C" {

const char *dpe_version(void) { return "deeperwin_b200 0.1 (sm_100a)"; }
const char *dpe_last_error(void) { return g_err; }

int dpe_model_create(const dpe_dims *dims, dpe_model **out) {
    if (!dims || !out) return set_error(DPE_ERR_ARG, "model_create: null argument");
    int e = validate_dims(*dims);
    if (e) return e;
    dpe_model *m = new (std::nothrow) dpe_model();
    if (!m) return set_error(DPE_ERR_ARG, "out of host memory");
    memset(m, 0, sizeof(*m));
    m->prof = new std::vector<dpe_model::ProfRec>();
    m->dims = *dims;
    build_leaves(m);
    m->derived_floats = derived_floats(*dims);
    cudaError_t ce = cudaMalloc(&m->params, (size_t)m->n_params * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->derived, m->derived_floats * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->R_dev, (size_t)dims->n_ion * 3 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->Z_dev, (size_t)dims->n_ion * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->eii_dev, sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->geom_flag_dev, sizeof(int32_t));
    if (ce == cudaSuccess) ce = cudaMemset(m->geom_flag_dev, 0, sizeof(int32_t));
    if (ce == cudaSuccess) ce = cudaMallocHost(&m->geom_pin, (size_t)DPE_GEOM_SLOTS * (4 * dims->n_ion + 1) * sizeof(float));
    for (int k = 0; k < DPE_GEOM_SLOTS && ce == cudaSuccess; ++k) ce = cudaEventCreateWithFlags(&m->geom_ev[k], cudaEventDisableTiming);
    m->det_flags = (getenv("DPE_DET_GENERIC") ? 1 : 0) | (getenv("DPE_DET_SIMT") ? 2 : 0);      // debug defaults; dpe_set_det_path overrides
    if (dims->use_taos) {
        const size_t gc = (size_t)dims->n_ion * dims->n_dets * dims->n_el;
        if (ce == cudaSuccess) ce = cudaMalloc(&m->tao_w, (size_t)dims->n_hidden_one_el[dims->n_iterations - 1] * gc * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc(&m->tao_ex[0], 2 * gc * sizeof(float));
        if (ce == cudaSuccess) m->tao_ex[1] = m->tao_ex[0] + gc;
    }
    if (ce != cudaSuccess) {
        int rc = check_cuda(ce, "cudaMalloc(model)");
        dpe_model_destroy(m);
        return rc;
    }
    bind_views(m);
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&m->n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || m->n_sm <= 0) m->n_sm = 148;
    }
    // tensor-core path: the big dense layers (h_el main blocks and the backflow matrices)
    {
        int e = DPE_OK;
        for (int it = 0; it < dims->n_iterations && !e; ++it) {
            e = tc_register_weight(m, m->it[it].w_main, m->it[it].k_main, m->it[it].d_out);
            if (!e) e = tc_register_weight(m, m->it[it].w_mean, 2 * m->it[it].d_in, m->it[it].d_out);
            if (!e && m->it[it].d_in <= 320) e = tc_register_weight(m, m->it[it].h_map.w, m->it[it].d_in, dims->emb_dim);
            if (!e) e = tc_register_weight(m, m->it[it].w_main, m->it[it].d_out, m->it[it].k_main, true);       // gradient pass: dx = dz W^T
        }
        const int cols = dims->n_dets * dims->n_el, dl = dims->n_hidden_one_el[dims->n_iterations - 1];
        if (dims->use_taos) { if (!e) e = tc_register_weight(m, m->tao_w, dl, dims->n_ion * cols); }
        else for (int sp = 0; sp < 2 && !e; ++sp) {
            e = tc_register_weight(m, m->bf_w[sp], dl, cols);
            if (!e) e = tc_register_weight(m, m->bf_w[sp], cols, dl, true);
        }
        if (e) { tc_destroy(m); cudaGetLastError(); }      // no tensor-core path: dense layers stay on the FP32 SIMT GEMM
        m->gemm_path = m->tc ? 1 : 0;
        { const char *g = getenv("DPE_MCMC_GRAPH"); m->mcmc_graph_mode = (g && g[0] == '0') ? 0 : 1; }
    }
    *out = m;
    return DPE_OK;
}

void dpe_model_destroy(dpe_model *m) {
    if (!m) return;
    cudaFree(m->params); cudaFree(m->derived); cudaFree(m->R_dev); cudaFree(m->Z_dev);
    cudaFree(m->eii_dev); cudaFree(m->geom_flag_dev);
    if (m->geom_pin) cudaFreeHost(m->geom_pin);
    for (int k = 0; k < DPE_GEOM_SLOTS; ++k)
        if (m->geom_ev[k]) cudaEventDestroy(m->geom_ev[k]);
    cudaFree(m->tao_w); cudaFree(m->tao_ex[0]);
    tc_destroy(m);
    mcmc_graphs_destroy(m);
    if (m->prof) {
        for (auto &r : *m->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        delete m->prof;
    }
    delete m;
}

int64_t dpe_param_count(const dpe_model *m) { return m ? m->n_params : 0; }
int32_t dpe_param_leaf_count(const dpe_model *m) { return m ? m->n_leaves : 0; }

int dpe_param_leaf(const dpe_model *m, int32_t leaf, int64_t *offset, int64_t *size, int32_t *rows, int32_t *cols) {
    if (!m || leaf < 0 || leaf >= m->n_leaves) return set_error(DPE_ERR_ARG, "param_leaf: bad index");
    if (offset) *offset = m->leaves[leaf].off;
    if (size) *size = m->leaves[leaf].size;
    if (rows) *rows = m->leaves[leaf].rows;
    if (cols) *cols = m->leaves[leaf].cols;
    return DPE_OK;
}

int dpe_model_set_params(dpe_model *m, const float *params_dev, int64_t n, void *stream) {
    if (!m || !params_dev) return set_error(DPE_ERR_ARG, "set_params: null argument");
    if (n != m->n_params) return set_error(DPE_ERR_ARG, "set_params: got %lld values, model has %lld", (long long)n, (long long)m->n_params);
    cudaStream_t s = (cudaStream_t)stream;
    DPE_CUDA(cudaMemcpyAsync(m->params, params_dev, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
    int e = launch_prepare_params(m, s);
    if (e) return e;
    if ((e = tc_refresh_weights(m, s))) return e;
    m->params_set = true;
    if (m->geom_set) return launch_prepare_geometry(m, s);   // him depends on the weights too
    return DPE_OK;
}

int dpe_model_set_geometry(dpe_model *m, const float *R_host, const int32_t *Z_host, void *stream) {
    if (!m || !R_host || !Z_host) return set_error(DPE_ERR_ARG, "set_geometry: null argument");
    const dpe_dims &d = m->dims;
    cudaStream_t s = (cudaStream_t)stream;
    // The host arrays may be temporaries: they are copied into a slot of a pinned staging ring and travel from there with
    // asynchronous copies, so the call neither synchronises the stream nor blocks on the transfer (the weight-sharing loop
    // changes the geometry every step, variational_optimization.py:356-387).  A slot is reused after DPE_GEOM_SLOTS calls; its
    // event tells whether the copy that last read it has finished.
    const int slot = m->geom_slot;
    m->geom_slot = (slot + 1) % DPE_GEOM_SLOTS;
    DPE_CUDA(cudaEventSynchronize(m->geom_ev[slot]));
    float *pin = m->geom_pin + (size_t)slot * (4 * d.n_ion + 1);
    float *Rp = pin, *Zf = pin + 3 * d.n_ion, *eiip = pin + 4 * d.n_ion;
    for (int J = 0; J < d.n_ion; ++J) {
        if (Z_host[J] < d.z_min || Z_host[J] > d.z_max) return set_error(DPE_ERR_ARG, "Z[%d]=%d outside [z_min, z_max]", J, Z_host[J]);
        Zf[J] = (float)Z_host[J];
    }
    memcpy(Rp, R_host, (size_t)d.n_ion * 3 * sizeof(float));
    // ion-ion repulsion (hamiltonian.py:25-31), float32 like the reference
    float eii = 0.f;
    for (int I = 0; I < d.n_ion; ++I)
        for (int J = I + 1; J < d.n_ion; ++J) {
            float dx = R_host[I * 3] - R_host[J * 3], dy = R_host[I * 3 + 1] - R_host[J * 3 + 1], dz = R_host[I * 3 + 2] - R_host[J * 3 + 2];
            eii += Zf[I] * Zf[J] / sqrtf(dx * dx + dy * dy + dz * dz);
        }
    *eiip = eii;
    DPE_CUDA(cudaMemcpyAsync(m->R_dev, Rp, (size_t)d.n_ion * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    DPE_CUDA(cudaMemcpyAsync(m->Z_dev, Zf, (size_t)d.n_ion * sizeof(float), cudaMemcpyHostToDevice, s));
    DPE_CUDA(cudaMemcpyAsync(m->eii_dev, eiip, sizeof(float), cudaMemcpyHostToDevice, s));
    DPE_CUDA(cudaEventRecord(m->geom_ev[slot], s));
    m->geom_set = true;
    if (m->params_set) return launch_prepare_geometry(m, s);
    return DPE_OK;
}

int dpe_model_set_geometry_dev(dpe_model *m, const float *R_dev, const int32_t *Z_dev, void *stream) {
    if (!m || !R_dev || !Z_dev) return set_error(DPE_ERR_ARG, "set_geometry_dev: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    int e = launch_geometry_from_device(m, R_dev, Z_dev, s);
    if (e) return e;
    m->geom_set = true;
    if (m->params_set) return launch_prepare_geometry(m, s);
    return DPE_OK;
}

int dpe_model_geometry_status(dpe_model *m, void *stream) {
    if (!m) return set_error(DPE_ERR_ARG, "geometry_status: null model");
    int32_t flag = 0;
    DPE_CUDA(cudaMemcpyAsync(&flag, m->geom_flag_dev, sizeof(flag), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DPE_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) return set_error(DPE_ERR_ARG, "set_geometry_dev: a nuclear charge was outside [z_min, z_max] (clamped)");
    return DPE_OK;
}

int dpe_model_set_tao_cache(dpe_model *m, const float *backflows_up_dev, const float *backflows_dn_dev,
                            const float *exponents_up_dev, const float *exponents_dn_dev, void *stream) {
    if (!m || !backflows_up_dev || !backflows_dn_dev || !exponents_up_dev || !exponents_dn_dev)
        return set_error(DPE_ERR_ARG, "set_tao_cache: null argument");
    if (!m->dims.use_taos) return set_error(DPE_ERR_STATE, "set_tao_cache: the model was created with envelope orbitals (use_taos = 0)");
    cudaStream_t s = (cudaStream_t)stream;
    int e = launch_tao_pack(m, backflows_up_dev, backflows_dn_dev, exponents_up_dev, exponents_dn_dev, s);
    if (e) return e;
    if ((e = tc_refresh_weights(m, s))) return e;
    m->tao_set = true;
    return DPE_OK;
}

size_t dpe_workspace_bytes(const dpe_model *m, int32_t n_walkers, int32_t mode) {
    if (!m || n_walkers <= 0) return 0;
    WsLayout L;
    plan(m->dims, n_walkers, mode == DPE_MODE_LAPLACIAN ? 3 * m->dims.n_el + 2 : 1, L);
    plan_mcmc(m->dims, n_walkers, L);
    return L.total_chunk + L.total_mcmc + 256;
}

int dpe_log_psi_sqr(dpe_model *m, const float *r_dev, int32_t n_walkers, float *phase_dev, float *log_psi_sqr_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!m || !r_dev || !log_psi_sqr_dev || !workspace_dev || n_walkers <= 0) return set_error(DPE_ERR_ARG, "log_psi_sqr: bad argument");
    return run_batched(m, r_dev, n_walkers, 1, (char *)workspace_dev, workspace_bytes, phase_dev, log_psi_sqr_dev, nullptr, nullptr,
                       nullptr, nullptr, (cudaStream_t)stream);
}

int dpe_local_energy(dpe_model *m, const float *r_dev, int32_t n_walkers, float *e_loc_dev, float *log_psi_sqr_dev,
                     float *grad_dev, float *e_kin_dev, float *e_pot_dev, void *workspace_dev, size_t workspace_bytes,
                     void *stream) {
    if (!m || !r_dev || !e_loc_dev || !workspace_dev || n_walkers <= 0) return set_error(DPE_ERR_ARG, "local_energy: bad argument");
    return run_batched(m, r_dev, n_walkers, 3 * m->dims.n_el + 2, (char *)workspace_dev, workspace_bytes, nullptr, log_psi_sqr_dev,
                       grad_dev, e_kin_dev, e_loc_dev, e_pot_dev, (cudaStream_t)stream);
}

// ---- Metropolis steps ------------------------------------------------------------------------------------------------------------
// One step = propose, network forward pass (~30 small launches), accept (+ controller).  A call that REPEATS an earlier call exactly
// (same state / count / workspace pointers, sizes, config and kernel-path settings -- the inter-step loop of a run) is replayed from a CUDA
// graph captured on its second occurrence: every launch argument is then identical by construction (the step number, step size and RNG keys
// live in device memory), and the per-launch CPU + front-end cost disappears.  First occurrences and non-repeating calls launch eagerly.
}  // extern "C"
namespace dpe {
struct McmcKey {
    dpe_mcmc_state st; dpe_mcmc_config cfg;
    int32_t B, n_steps, recompute, run_controller, gemm_path, det_flags;
    void *counts, *ws; size_t ws_bytes;
};
struct McmcGraph { McmcKey key; cudaGraphExec_t exec; int64_t launches; uint64_t used; };
struct McmcGraphCache { std::vector<McmcKey> seen; std::vector<McmcGraph> graphs; uint64_t tick = 0; cudaStream_t capture_stream = nullptr; };      // `seen`: the last few eager calls
void mcmc_graphs_destroy(dpe_model *m) {
    if (!m->mcmc_graphs) return;
    for (auto &g : m->mcmc_graphs->graphs) cudaGraphExecDestroy(g.exec);
    if (m->mcmc_graphs->capture_stream) cudaStreamDestroy(m->mcmc_graphs->capture_stream);
    delete m->mcmc_graphs;
    m->mcmc_graphs = nullptr;
}

static int mcmc_steps_eager(dpe_model *m, const dpe_mcmc_state *st, int32_t B, int32_t n_steps, const dpe_mcmc_config *cfg, int32_t recompute_log_psi,
                            int32_t run_controller, int32_t *accept_counts_dev, char *ws, size_t workspace_bytes, cudaStream_t s) {
    WsLayout L;
    plan_mcmc(m->dims, B, L);
    float *r_prop = (float *)(ws + L.r_prop), *lp_prop = (float *)(ws + L.lp_prop), *thr = (float *)(ws + L.thr);
    uint32_t *new_keys = (uint32_t *)(ws + L.new_keys);
    float *log_q = cfg->proposal >= 3 ? (float *)(ws + L.log_q) : nullptr;
    char *ws_net = ws + L.total_mcmc;
    size_t ws_net_bytes = workspace_bytes - L.total_mcmc;
    int e;
    if (recompute_log_psi)
        if ((e = run_batched(m, st->r_dev, B, 1, ws_net, ws_net_bytes, nullptr, st->log_psi_sqr_dev, nullptr, nullptr, nullptr, nullptr, s))) return e;
    if (n_steps > 0) DPE_CUDA(cudaMemsetAsync(accept_counts_dev, 0, (size_t)n_steps * sizeof(int32_t), s));
    for (int t = 0; t < n_steps; ++t) {
        // without the in-call controller step_nr is advanced afterwards (dpe_mcmc_controller): offset the moved electron by t
        if ((e = launch_propose(m, st, B, *cfg, run_controller ? 0 : t, r_prop, thr, new_keys, log_q, s))) return e;
        m->launches++;
        if ((e = run_batched(m, r_prop, B, 1, ws_net, ws_net_bytes, nullptr, lp_prop, nullptr, nullptr, nullptr, nullptr, s))) return e;
        if ((e = launch_accept(st, B, m->dims.n_el, r_prop, lp_prop, thr, new_keys, log_q, cfg->max_age, nullptr, accept_counts_dev + t, s))) return e;
        m->launches++;
        if (run_controller) {
            if ((e = launch_controller(st, accept_counts_dev + t, 1, B, *cfg, s))) return e;
            m->launches++;
        }
    }
    return DPE_OK;
}
}  // namespace dpe
extern "C" {

int dpe_mcmc_steps(dpe_model *m, const dpe_mcmc_state *st, int32_t B, int32_t n_steps, const dpe_mcmc_config *cfg,
                   int32_t recompute_log_psi, int32_t run_controller, int32_t *accept_counts_dev, void *workspace_dev,
                   size_t workspace_bytes, void *stream) {
    if (!m || !st || !cfg || !workspace_dev || B <= 0 || n_steps < 0) return set_error(DPE_ERR_ARG, "mcmc_steps: bad argument");
    if (!st->r_dev || !st->log_psi_sqr_dev || !st->walker_age_dev || !st->rng_state_dev || !st->stepsize_dev || !st->step_nr_dev || !st->acc_rate_dev)
        return set_error(DPE_ERR_ARG, "mcmc_steps: state has null fields");
    if (n_steps > 0 && !accept_counts_dev) return set_error(DPE_ERR_ARG, "mcmc_steps: accept_counts_dev is null");
    if (cfg->proposal < 0 || cfg->proposal > 5) return set_error(DPE_ERR_UNSUPPORTED, "mcmc_steps: proposal %d (0 normal, 1 cauchy, 2 normal_one_el, 3 local, 4 local_one_el, 5 langevin)", cfg->proposal);
    if (cfg->proposal >= 3 && !(cfg->r_min > 0.f && cfg->r_max >= cfg->r_min)) return set_error(DPE_ERR_ARG, "mcmc_steps: local proposals need 0 < r_min <= r_max");
    if (cfg->proposal >= 3 && !m->geom_set) return set_error(DPE_ERR_STATE, "mcmc_steps: set_geometry must be called first");
    cudaStream_t s = (cudaStream_t)stream;
    WsLayout L;
    plan_mcmc(m->dims, B, L);
    if (workspace_bytes <= L.total_mcmc) return set_error(DPE_ERR_WORKSPACE, "workspace too small for the Metropolis scratch");
    char *ws = (char *)workspace_dev;
    const bool try_graph = m->mcmc_graph_mode == 1 && !m->profile && n_steps > 0;
    if (!try_graph) return mcmc_steps_eager(m, st, B, n_steps, cfg, recompute_log_psi, run_controller, accept_counts_dev, ws, workspace_bytes, s);

    McmcKey key;
    memset(&key, 0, sizeof(key));
    key.st = *st; key.cfg = *cfg; key.B = B; key.n_steps = n_steps; key.recompute = recompute_log_psi; key.run_controller = run_controller;
    key.gemm_path = m->gemm_path; key.det_flags = m->det_flags; key.counts = accept_counts_dev; key.ws = ws; key.ws_bytes = workspace_bytes;
    if (!m->mcmc_graphs) m->mcmc_graphs = new McmcGraphCache();
    McmcGraphCache &gc = *m->mcmc_graphs;
    for (auto &g : gc.graphs)
        if (!memcmp(&g.key, &key, sizeof(key))) {
            DPE_CUDA(cudaGraphLaunch(g.exec, s));
            g.used = ++gc.tick;
            m->launches += g.launches;
            return DPE_OK;
        }
    bool repeat = false;
    for (size_t k = 0; k < gc.seen.size() && !repeat; ++k)
        if (!memcmp(&gc.seen[k], &key, sizeof(key))) { repeat = true; gc.seen.erase(gc.seen.begin() + k); }
    if (!repeat) {
        if (gc.seen.size() >= 64) gc.seen.erase(gc.seen.begin());
        gc.seen.push_back(key);
    }
    if (!repeat) return mcmc_steps_eager(m, st, B, n_steps, cfg, recompute_log_psi, run_controller, accept_counts_dev, ws, workspace_bytes, s);

    // second occurrence: capture.  Any failure ends the capture, clears the error and runs the call eagerly.
    const int64_t launches0 = m->launches;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    // captured on a private stream (the caller's may be the legacy default stream, which cannot be captured); the graph is launched on the caller's
    if (!gc.capture_stream && cudaStreamCreateWithFlags(&gc.capture_stream, cudaStreamNonBlocking) != cudaSuccess) gc.capture_stream = nullptr;
    bool ok = gc.capture_stream && cudaStreamBeginCapture(gc.capture_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
        const int e = mcmc_steps_eager(m, st, B, n_steps, cfg, recompute_log_psi, run_controller, accept_counts_dev, ws, workspace_bytes, gc.capture_stream);
        const cudaError_t ce = cudaStreamEndCapture(gc.capture_stream, &graph);
        ok = e == DPE_OK && ce == cudaSuccess && graph != nullptr;
    }
    const int64_t captured = m->launches - launches0;
    m->launches = launches0;
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        m->mcmc_graph_mode = 0;                    // do not try again on this model
        return mcmc_steps_eager(m, st, B, n_steps, cfg, recompute_log_psi, run_controller, accept_counts_dev, ws, workspace_bytes, s);
    }
    if (gc.graphs.size() >= 48) {                  // bounded cache (a weight-sharing run keeps two state buffers per geometry): drop the least recently used
        size_t lru = 0;
        for (size_t k = 1; k < gc.graphs.size(); ++k)
            if (gc.graphs[k].used < gc.graphs[lru].used) lru = k;
        cudaGraphExecDestroy(gc.graphs[lru].exec);
        gc.graphs.erase(gc.graphs.begin() + lru);
    }
    gc.graphs.push_back(McmcGraph{key, exec, captured, ++gc.tick});
    DPE_CUDA(cudaGraphLaunch(exec, s));
    m->launches += captured;
    return DPE_OK;
}

int dpe_set_mcmc_graph(dpe_model *m, int32_t mode) {
    if (!m || (mode != 0 && mode != 1)) return set_error(DPE_ERR_ARG, "set_mcmc_graph: mode must be 0 (eager) or 1 (graph replay of repeated calls)");
    m->mcmc_graph_mode = mode;
    if (!mode) mcmc_graphs_destroy(m);
    return DPE_OK;
}
int dpe_get_mcmc_graph(const dpe_model *m) { return m ? m->mcmc_graph_mode : -1; }

int64_t dpe_debug_ws_offset(const dpe_model *m, int32_t n_walkers, int32_t mode, const char *name) {
    if (!m || !name || n_walkers <= 0) return -1;
    WsLayout L;
    plan(m->dims, n_walkers, mode == DPE_MODE_LAPLACIAN ? 3 * m->dims.n_el + 2 : 1, L);
    if (!strcmp(name, "x0")) return (int64_t)L.x[0];
    if (!strcmp(name, "x1")) return (int64_t)L.x[1];
    if (!strcmp(name, "hm")) return (int64_t)L.hm;
    if (!strcmp(name, "mean")) return (int64_t)L.mean;
    if (!strcmp(name, "add")) return (int64_t)L.add;
    if (!strcmp(name, "mo")) return (int64_t)L.mo;
    if (!strcmp(name, "det")) return (int64_t)L.det;
    if (!strcmp(name, "epot")) return (int64_t)L.epot;
    if (!strcmp(name, "ldx")) return (int64_t)L.ldx;
    if (!strncmp(name, "pw", 2) && name[2] >= '0' && name[2] < '0' + DPE_MAX_ITER) return (int64_t)L.pw_it[name[2] - '0'];
    if (!strncmp(name, "ei", 2) && name[2] >= '0' && name[2] < '0' + DPE_MAX_ITER) return (int64_t)L.ei_it[name[2] - '0'];
    return -1;
}

int dpe_set_gemm_path(dpe_model *m, int32_t path) {
    if (!m || (path != 0 && path != 1)) return set_error(DPE_ERR_ARG, "set_gemm_path: path must be 0 or 1");
    if (path == 1 && !m->tc) return set_error(DPE_ERR_UNSUPPORTED, "tensor-core path unavailable (TMA descriptor encode failed)");
    m->gemm_path = path;
    return DPE_OK;
}

int dpe_debug_gemm(dpe_model *m, int32_t path, const float *a_dev, int32_t lda, const float *w_dev, float *c_dev, int32_t ldc,
                   int32_t M, int32_t N, int32_t K, int32_t seg_len, int32_t a_seg_stride, int32_t a_seg_off, int32_t c_seg_stride,
                   int32_t c_seg_off, int32_t c_col_off, void *stream) {
    if (!m || !a_dev || !w_dev || !c_dev) return set_error(DPE_ERR_ARG, "debug_gemm: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    GemmArgs g;
    g.A = a_dev; g.lda = lda; g.W = w_dev; g.ldw = N; g.C = c_dev; g.ldc = ldc;
    g.a_seg_len = g.c_seg_len = seg_len > 0 ? seg_len : M;
    g.a_seg_stride = a_seg_stride; g.a_seg_off = a_seg_off; g.c_seg_stride = c_seg_stride; g.c_seg_off = c_seg_off; g.c_col_off = c_col_off;
    g.M = M; g.N = N; g.K = K;
    if (path == 0) return launch_gemm_simt(m, g, s);
    int e = tc_register_weight(m, w_dev, K, N);
    if (e) return e;
    if ((e = tc_refresh_weights(m, s))) return e;
    e = launch_gemm_tc(m, g, s);
    if (e == DPE_ERR_UNSUPPORTED) return set_error(e, "debug_gemm: shape not supported by the tensor-core kernel");
    return e;
}
int dpe_get_gemm_path(const dpe_model *m) { return m ? m->gemm_path : -1; }
int dpe_set_det_path(dpe_model *m, int32_t flags) {
    if (!m || flags < 0 || flags > 3) return set_error(DPE_ERR_ARG, "set_det_path: flags must be in [0, 3]");
    m->det_flags = flags;
    return DPE_OK;
}
int dpe_get_det_path(const dpe_model *m) { return m ? m->det_flags : -1; }
int dpe_profile_enable(dpe_model *m, int32_t on) {
    if (!m) return set_error(DPE_ERR_ARG, "profile_enable: null model");
    m->profile = on != 0;
    return DPE_OK;
}

int dpe_profile_collect(dpe_model *m, int32_t klass, double *ms, int64_t *count, double *flops) {
    if (!m || !ms || !count || !flops) return set_error(DPE_ERR_ARG, "profile_collect: null argument");
    DPE_CUDA(cudaDeviceSynchronize());
    *ms = 0.0; *count = 0; *flops = 0.0;
    std::vector<dpe_model::ProfRec> keep;
    for (auto &r : *m->prof) {
        if (r.klass != klass) { keep.push_back(r); continue; }
        float t = 0.f;
        DPE_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        *ms += t; *count += 1; *flops += r.flops;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    m->prof->swap(keep);
    return DPE_OK;
}

int dpe_profile_launches(dpe_model *m, int32_t klass, double *ms_arr, double *flops_arr, int32_t cap, int32_t *n) {
    if (!m || !ms_arr || !flops_arr || !n) return set_error(DPE_ERR_ARG, "profile_launches: null argument");
    DPE_CUDA(cudaDeviceSynchronize());
    *n = 0;
    for (auto &r : *m->prof) {
        if (r.klass != klass || *n >= cap) continue;
        float t = 0.f;
        DPE_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms_arr[*n] = t; flops_arr[*n] = r.flops; ++*n;
    }
    return DPE_OK;
}

int dpe_profile_stages(dpe_model *m, double *ms_arr, int64_t *count_arr, int32_t cap) {
    if (!m || !ms_arr || !count_arr || cap < ST_COUNT) return set_error(DPE_ERR_ARG, "profile_stages: need room for %d stages", (int)ST_COUNT);
    DPE_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < cap; ++k) { ms_arr[k] = 0.0; count_arr[k] = 0; }
    std::vector<dpe_model::ProfRec> keep;
    for (auto &r : *m->prof) {
        if (r.stage < 0) { keep.push_back(r); continue; }
        float t = 0.f;
        DPE_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms_arr[r.stage] += t; count_arr[r.stage] += 1;
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
    }
    m->prof->swap(keep);
    return ST_COUNT;
}

int64_t dpe_launch_count(const dpe_model *m) { return m ? m->launches : 0; }

}  // extern "C"
