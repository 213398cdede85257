// api.cu -- C ABI entry points for the ISS kernels (see include/fruits_b200.h).
#include <stdarg.h>
#include <string.h>

#include "lns_inst.cuh"

namespace fb {

char *err_buf()
{
    static thread_local char buf[512] = {0};
    return buf;
}

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

static int fill_params(LnsParams &p, const fb_iss_plan *plan, const fb_batch *b, int rmax,
                       long long t_max = 65535)
{
    FB_REQUIRE(plan && b, "null plan or batch");
    FB_REQUIRE(plan->n_rows >= 1 && plan->n_rows <= rmax,
               "plan uses %d rows per block, kernel supports %d", plan->n_rows, rmax);
    FB_REQUIRE(plan->n_used_dims >= 1 && plan->n_used_dims <= FB_MAX_USED_DIMS,
               "n_used_dims=%d out of range", plan->n_used_dims);
    FB_REQUIRE(plan->n_alphas >= 0 && plan->n_alphas <= FB_MAX_ALPHAS, "n_alphas=%d out of range",
               plan->n_alphas);
    FB_REQUIRE(plan->max_depth >= 1 && plan->max_depth <= FB_RING - 64,
               "word length %d not supported (max %d)", plan->max_depth, FB_RING - 64);
    // (the sieving policies pack two 16-bit counters per register)
    FB_REQUIRE(b->t >= 1 && b->t <= t_max, "series length %lld not supported (1..%lld)",
               (long long)b->t, t_max);
    FB_REQUIRE(b->n >= 0 && b->d >= 1, "bad batch shape");
    FB_REQUIRE(plan->weight_mode == FB_WEIGHT_NONE || b->g != nullptr,
               "weighted ISS needs a lookup table");
    memset(&p, 0, sizeof(p));
    p.slots = plan->slots;
    p.row_pub = plan->row_pub;
    p.row_weight = plan->row_weight;
    p.X = b->X;
    p.g = b->g;
    p.stats = b->stats;
    p.n = b->n; p.d = b->d; p.t = b->t; p.g_ld = b->g_ld;
    p.n_blocks = plan->n_blocks; p.n_rows = plan->n_rows; p.n_emit = plan->n_emit;
    p.du = plan->n_used_dims;
    p.na = plan->weight_mode == FB_WEIGHT_NONE ? 0 : (plan->n_alphas > 0 ? plan->n_alphas : 1);
    p.max_depth = plan->max_depth;
    for (int a = 0; a < FB_MAX_ALPHAS; a++) p.alphas[a] = plan->alphas[a];
    for (int u = 0; u < plan->n_used_dims; u++) {
        p.dims[u] = plan->dims[u];
        FB_REQUIRE(p.dims[u].raw_dim >= 0 && p.dims[u].raw_dim < b->d,
                   "plan reads dimension %d but the input has %lld", p.dims[u].raw_dim + 1,
                   (long long)b->d);
        FB_REQUIRE(!p.dims[u].std || b->stats, "standardised dimension needs batch.stats");
        p.any_inc |= p.dims[u].inc;
        p.any_std |= p.dims[u].std;
    }
    return 0;
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_abi_version(void) { return FB_ABI_VERSION; }

const char *fb_last_error(void) { return err_buf(); }

int fb_device_info(int *sm_count, int *cc_major, int *cc_minor, int *smem_optin)
{
    int dev = 0;
    FB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    FB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin) *smem_optin = (int)prop.sharedMemPerBlockOptin;
    return 0;
}

int fb_iss_materialize(const fb_iss_plan *plan, const fb_batch *batch, double *out, void *stream)
{
    LnsParams p;
    int rc = fill_params(p, plan, batch, RMAX_MAT, (1LL << 30));
    if (rc) return rc;
    FB_REQUIRE(out != nullptr, "null output");
    p.out = out;
    p.mat_stage = plan->n_emit <= LNS_STAGE_ROWS;      // few rows: staged 256-byte row writes
    return lns_run_mat(p, plan->semiring, plan->weight_mode, (cudaStream_t)stream);
}

int fb_slice_rows(int policy)
{
    switch (policy) {
    case POL_MAT: return RMAX_MAT;
    case POL_A: return RMAX_A;
    case POL_D: return RMAX_D;
    case POL_P: return RMAX_P;
    case POL_M: return RMAX_M;
    case POL_G: return RMAX_G;
    default: return -1;
    }
}

/* Smallest fused policy that covers the sieve plan; < 0 if none does. */
int fb_slice_policy(const fb_sieve_plan *sv, int bounded_hi, int bounded_mm)
{
    if (!sv || sv->n_feats < 1 || sv->n_feats > FB_MAX_FEATS) return FB_EINVAL;
    bool c[3] = {false, false, false}, s[3] = {false, false, false};
    bool ppv = false, mx = false, mn = false;
    for (int f = 0; f < sv->n_feats; f++) {
        const int k = sv->kind[f], a = sv->arg[f];
        if (k == FB_FEAT_CNT || k == FB_FEAT_AVG) {
            if (a < 0 || a > 2) return FB_ENOSUP;
            c[a] = true;
            if (k == FB_FEAT_AVG) s[a] = true;
        } else if (k == FB_FEAT_PPV) ppv = true;
        else if (k == FB_FEAT_MAX) mx = true;
        else if (k == FB_FEAT_MIN) mn = true;
        else if (k >= FB_FEAT_XPI && k <= FB_FEAT_CPV) return FB_ENOSUP;   // generated kernels only
        else if (k != FB_FEAT_END) return FB_EINVAL;
    }
    if (bounded_hi || bounded_mm) return POL_G;
    const bool only1 = !c[0] && !c[2];
    if (only1 && !ppv && !mx && !mn && !s[1]) return POL_A;
    if (only1 && !ppv && !mx && !mn) return POL_D;
    if (only1 && !s[1]) return POL_P;
    if (!ppv && !mx && !mn) return POL_M;
    return POL_G;
}

int fb_slice_features_ex(const fb_iss_plan *plan, const fb_batch *batch, const fb_sieve_plan *sv,
                         double *out, int64_t out_ld, int64_t col0, int policy, int sanitize,
                         void *stream)
{
    FB_REQUIRE(sv && out, "null sieve plan or output");
    FB_REQUIRE(sv->n_feats >= 1 && sv->n_feats <= FB_MAX_FEATS, "n_feats=%d out of range",
               sv->n_feats);
    FB_REQUIRE(sv->thresholds != nullptr, "null threshold table");
    const int rmax = fb_slice_rows(policy);
    FB_REQUIRE(rmax > 0 && policy != POL_MAT, "bad policy %d", policy);
    LnsParams p;
    int rc = fill_params(p, plan, batch, rmax);
    if (rc) return rc;
    p.out = out;
    p.out_ld = out_ld;
    p.col0 = col0;
    p.thr = sv->thresholds;
    p.n_feats = sv->n_feats;
    p.sanitize = sanitize;
    for (int f = 0; f < sv->n_feats; f++) {
        p.feat_kind[f] = sv->kind[f];
        p.feat_arg[f] = sv->arg[f];
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (policy) {
    case POL_A: return lns_run_a(p, plan->semiring, plan->weight_mode, st);
    case POL_D: return lns_run_d(p, plan->semiring, plan->weight_mode, st);
    case POL_P: return lns_run_p(p, plan->semiring, plan->weight_mode, st);
    case POL_M: return lns_run_m(p, plan->semiring, plan->weight_mode, st);
    default: return lns_run_g(p, plan->semiring, plan->weight_mode, st);
    }
}

int fb_slice_features(const fb_iss_plan *plan, const fb_batch *batch, const fb_sieve_plan *sv,
                      double *out, int64_t out_ld, int64_t col0, void *stream)
{
    const int policy = fb_slice_policy(sv, 1, 1);
    if (policy < 0) return set_err(policy, "sieve plan not supported");
    return fb_slice_features_ex(plan, batch, sv, out, out_ld, col0, policy, 0, stream);
}

}  // extern "C"
