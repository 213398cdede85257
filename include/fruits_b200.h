/*
 * fruits_b200.h -- C ABI of the B200-native FRUITS hot path.
 *
 * The reference (irkri/fruits 1.0.0) is pure Python + numba and has no FFI;
 * its boundary is the Python class API.  This header is the boundary a
 * maintainer of the reference would bind with ctypes:
 * every entry point replaces one numba kernel / numpy call of the reference,
 * cited as file:line relative to the reference checkout.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers (cudaMalloc'ed, 8-byte aligned)
 *     unless the name ends in `_h` (host pointer);
 *   - float64 arrays are C-ordered; `ld` arguments are row strides in elements;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous with
 *     respect to the host unless stated otherwise;
 *   - return value: 0 on success, a cudaError_t (> 0) for CUDA failures,
 *     a negative FB_E* code for argument errors; fb_last_error() returns a
 *     thread-local message.  There is no CPU fallback anywhere.
 */
#ifndef FRUITS_B200_H
#define FRUITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 1

#define FB_EINVAL  (-1)   /* bad argument */
#define FB_ENOSUP  (-2)   /* configuration not supported by this build */

/* ---- enums --------------------------------------------------------- */
/* fruits/iss/semiring.py: Reals (:161), Arctic (:341) */
#define FB_SEMIRING_REALS   0
#define FB_SEMIRING_ARCTIC  1
#define FB_SEMIRING_BAYESIAN 2  /* fb_bayes_word and generated kernels; not the generic trie kernel */
/* fruits/iss/semiring.py:27-35: no weighting, Weighting.total True/False */
#define FB_WEIGHT_NONE      0
#define FB_WEIGHT_TOTAL     1
#define FB_WEIGHT_NONTOTAL  2

/* Maximum number of rows (x32 lanes) one warp owns in the ISS kernel. */
#define FB_MAX_ROWS 16
/* Maximum number of distinct input dimensions one ISS plan may reference. */
#define FB_MAX_USED_DIMS 7
/* Time ring length of the ISS kernel; arctic word length must stay below
 * FB_RING - 32. */
#define FB_RING 256
/* Maximum number of distinct alpha values in one weighted ISS plan. */
#define FB_MAX_ALPHAS 4

/* One trie node placed on one lane of one row of a warp ("slot").
 * The host compiles ISS words into a prefix trie (fruits/iss/cache.py:17-37
 * defines which prefixes are emitted) and lays the nodes out as
 * blocks[n_blocks][FB rows][32 lanes]. */
typedef struct fb_slot {
    /* Reals: up to 16 letter occurrences, 4 bits each, applied in order:
     *   bits[2:0] = used-dimension index (7 = constant one / padding),
     *   bit 3     = 1: divide instead of multiply (negative exponent).
     * Arctic: up to 8 (dimension, exponent) pairs, 8 bits each:
     *   bits[2:0] = used-dimension index, bits[7:3] = signed exponent. */
    uint32_t letter_lo;
    uint32_t letter_hi;
    int16_t parent;   /* slot index (row*32+lane) of the parent in this block, -1 = root */
    int16_t emit;     /* index of the emitted iterated sum, -1 = not emitted */
    uint8_t depth;    /* 1-based position of the letter in its word */
    uint8_t aidx;     /* index into alphas[] of this level's alpha */
    uint8_t weight;   /* number of occurrences / pairs used */
    uint8_t flags;    /* bit0: slot valid, bit1: has children in this block */
} fb_slot;

/* How one used dimension is produced from the raw input on load
 * (fruits/preparation/transform.py:15-89 INC, :92-158 STD,
 *  fruits/preparation/wrapper.py:53-103 NEW). */
typedef struct fb_dim {
    int32_t raw_dim;   /* dimension of X to read */
    int32_t inc;       /* 1: x[t]-x[t-1] with zero padding (INC(1,1,True)) */
    int32_t std;       /* 1: (v - mean)/(denom) with stats[n][u][2] */
    int32_t pad;
} fb_dim;

typedef struct fb_iss_plan {
    int32_t semiring;      /* FB_SEMIRING_* */
    int32_t weight_mode;   /* FB_WEIGHT_* */
    int32_t n_blocks;      /* warp-blocks per series */
    int32_t n_rows;        /* rows used per block (<= rows the kernel supports) */
    int32_t n_emit;        /* number of emitted iterated sums */
    int32_t n_used_dims;   /* <= FB_MAX_USED_DIMS */
    int32_t n_alphas;      /* <= FB_MAX_ALPHAS */
    int32_t max_depth;     /* longest word */
    float   alphas[FB_MAX_ALPHAS];
    fb_dim  dims[FB_MAX_USED_DIMS];
    const fb_slot *slots;        /* device: [n_blocks][n_rows][32] */
    const uint32_t *row_pub;     /* device: [n_blocks] bit j set: row j holds a node with children */
    const uint8_t *row_weight;   /* device: [n_blocks][n_rows] max letter weight in the row */
} fb_iss_plan;

/* Input batch: X[n][d][t] float64 (fruits/fruit.py:138-173 transform input),
 * optional lookup g (fruits/iss/weighting.py get_lookup): g_ld == 0 means one
 * shared row g[t] (Indices / Plateaus), else g[n][t] with row stride g_ld.
 * stats: [n][n_used_dims][2] = (mean, std + eps) for dims with std=1. */
typedef struct fb_batch {
    const double *X;
    int64_t n, d, t;
    const double *g;
    int64_t g_ld;
    const double *stats;
} fb_batch;

/* ---- sieve fusion --------------------------------------------------- */
/* Feature kinds evaluated in the epilogue of the fused kernel; every kind
 * follows one reference sieve (default cut = -1, i.e. the whole series):
 *   CNT: NPI  fruits/sieving/increment.py:101-129
 *   AVG: MPI  fruits/sieving/increment.py:132-163
 *   PPV:      fruits/sieving/implicit.py:114-129
 *   MAX/MIN:  fruits/sieving/segment.py:107-200
 *   END:      fruits/sieving/segment.py:203-225 */
#define FB_FEAT_CNT 0   /* arg = increment depth 0..2 */
#define FB_FEAT_AVG 1   /* arg = increment depth 0..2 */
#define FB_FEAT_PPV 2
#define FB_FEAT_MAX 3
#define FB_FEAT_MIN 4
#define FB_FEAT_END 5
/* generated kernels only (the generic kernel answers FB_ENOSUP):
 *   XPI: fruits/sieving/increment.py:166-199   LPI: :202-239   (arg = increment depth)
 *   CUR: fruits/sieving/segment.py:228-260 (also AVG / STD, which call its backend)
 *   CPV: fruits/sieving/implicit.py:156-190 */
#define FB_FEAT_XPI 6
#define FB_FEAT_LPI 7
#define FB_FEAT_CUR 8
#define FB_FEAT_CPV 9
#define FB_MAX_FEATS 16

/* Threshold table layout per emitted iterated sum (row of FB_NTHR doubles):
 *   [0..1] (lo, hi] of increment depth 0   [2..3] depth 1   [4..5] depth 2
 *   [6]    PPV threshold                   [7] CPV threshold
 *   [8..9] (lo, hi] of MAX                 [10..11] (lo, hi] of MIN
 *   [12..13] (lo, hi] of CUR               [14..15] unused */
#define FB_NTHR 16

typedef struct fb_sieve_plan {
    int32_t n_feats;                 /* features per emitted iterated sum */
    int32_t kind[FB_MAX_FEATS];      /* FB_FEAT_* */
    int32_t arg[FB_MAX_FEATS];       /* increment depth for CNT/AVG */
    const double *thresholds;        /* device: [n_emit][FB_NTHR] */
} fb_sieve_plan;

/* ---- entry points ---------------------------------------------------- */
#if defined(__GNUC__)
#define FB_API __attribute__((visibility("default")))
#else
#define FB_API
#endif

FB_API int fb_abi_version(void);
FB_API const char *fb_last_error(void);
/* sm count, compute capability and bytes of shared memory per block opt-in */
FB_API int fb_device_info(int *sm_count, int *cc_major, int *cc_minor, int *smem_optin);

/* -- ISS + sieves (the hot path) -- */

/* Fused kernel variants ("policies"): the set of per-node accumulators that
 * is compiled in.  fb_slice_policy() returns the smallest one covering a
 * sieve plan (bounded_hi: some (lo, hi] interval has a finite hi;
 * bounded_mm: MAX/MIN use an interval other than (-inf, +inf]).
 * fb_slice_rows(policy) is the number of rows of 32 trie nodes one warp owns
 * in that variant -- the block capacity the host must compile the plan for
 * (policy 0 is the materialising kernel of fb_iss_materialize). */
FB_API int fb_slice_policy(const fb_sieve_plan *sieves, int bounded_hi, int bounded_mm);
FB_API int fb_slice_rows(int policy);

/* Fused FruitSlice.transform (fruits/fruit.py:498-553): preparateur-on-load,
 * ISS over the prefix trie, sieves in registers.  Writes
 *   out[i*out_ld + col0 + emit*n_feats + f]   for i < n.
 * sanitize != 0 applies np.nan_to_num(nan=0) (fruits/fruit.py:172) on store.
 * The ISS tensor is never written to memory. */
FB_API int fb_slice_features_ex(const fb_iss_plan *plan, const fb_batch *batch,
                                const fb_sieve_plan *sieves, double *out, int64_t out_ld,
                                int64_t col0, int policy, int sanitize, void *stream);
/* Same with the most general policy (plan must be compiled for its rows). */
FB_API int fb_slice_features(const fb_iss_plan *plan, const fb_batch *batch,
                             const fb_sieve_plan *sieves, double *out, int64_t out_ld,
                             int64_t col0, void *stream);

/* -- plan-specialised kernels --
 * The fastest form of the fused FruitSlice.transform: the host compiles the
 * slice (word trie, semiring, weighting mode, sieve set) into CUDA source in
 * which every trie node is a register of a thread (one thread = one series x
 * one part of the trie; generator: fruits_b200/_jit.py).  The library
 * compiles that source for sm_100a with NVRTC (one translation unit per trie
 * part, in parallel), links the parts with nvJitLink, loads the cubin and
 * launches it.  Semantics are those of fb_slice_features_ex; the source
 * contract is
 *   extern "C" __global__ void fb_jit_slice(struct Args)   and
 *   __constant__ double TH[]    (compact threshold table, see _jit.py).  */
typedef struct fb_jit_kernel fb_jit_kernel;

typedef struct fb_jit_geometry {
    int32_t n_parts;         /* parts the trie was split into (multiple of parts_per_cta) */
    int32_t parts_per_cta;   /* warps of one CTA working on the same 32 series */
    int32_t groups_per_cta;  /* groups of 32 series per CTA */
    int32_t smem_bytes;      /* dynamic shared memory of the generated kernel */
} fb_jit_geometry;

/* CUDA C++ source -> sm_100a cubin (malloc'ed, release with fb_jit_free).
 * relocatable != 0: a translation unit to be linked with fb_jit_link (the
 * parts of a trie are compiled in parallel, thread-safe); max_registers > 0
 * caps the registers per thread.  log (may be NULL) receives the compiler
 * log.  Needs no GPU. */
FB_API int fb_jit_compile(const char *src, const char *name, int relocatable, int max_registers,
                          void **cubin, size_t *size, char *log, size_t log_cap);
/* Link relocatable cubins into one loadable cubin (nvJitLink). */
FB_API int fb_jit_link(const void *const *cubins, const size_t *sizes, int n, void **out,
                       size_t *size, char *log, size_t log_cap);
FB_API void fb_jit_free(void *p);
FB_API int fb_jit_load(const void *cubin, size_t size, fb_jit_kernel **out);
FB_API int fb_jit_unload(fb_jit_kernel *k);
/* Fused FruitSlice.transform (fruits/fruit.py:498-553) with a generated kernel.
 * extra: weighting rows [n or 1][rows][t] (Reals: exp(+a g), exp(-a g) per
 * alpha; Arctic: g), extra_ld = row stride between series in elements, 0 if
 * shared.  thr: device table of n_thr doubles copied into the kernel's
 * constant bank on `stream` before the launch.  Launches of one kernel
 * object must be issued on one stream at a time. */
FB_API int fb_jit_slice_features(fb_jit_kernel *k, const fb_jit_geometry *geo,
                                 const fb_batch *batch, const double *extra, int64_t extra_ld,
                                 const double *thr, int64_t n_thr, double *out, int64_t out_ld,
                                 int64_t col0, int sanitize, void *stream);
/* (`sanitize`: bit 0 = np.nan_to_num of the features (fruits/fruit.py:172), bit 1 =
 * `out` is an NVSwitch multicast mapping, stored to with multimem.st) */
/* The same with a cut: the segment sieves (NPI, MPI, XPI, LPI, MAX, MIN, CUR, END)
 * look at the time steps [0, cuts[series]) only -- one int or float ("coquantile")
 * `cut` argument of the reference's SegmentSieve (fruits/sieving/segment.py:51-64,
 * fruits/cache.py:16-22, :98-107); END is the value at cuts[series] - 1.  cuts =
 * DEVICE int32 [n] or NULL (whole series).  The kernel must have been generated
 * for a cut (fruits_b200/_jit.py, SieveSet.cut). */
FB_API int fb_jit_slice_features_cut(fb_jit_kernel *k, const fb_jit_geometry *geo,
                                     const fb_batch *batch, const double *extra,
                                     int64_t extra_ld, const double *thr, int64_t n_thr,
                                     const int32_t *cuts, double *out, int64_t out_ld,
                                     int64_t col0, int sanitize, void *stream);

/* Lane-per-node form for deep Arctic tries (generator: fruits_b200/_jit_chain.py):
 * the 24-48-letter alternating-sign chains of experiments/fruit_reduced.py:42-49
 * and fruit_general.py:42-51 (fruits/iss/semiring.py:314-338).  One warp owns one
 * series and one block of the trie, lanes are skewed in time by their depth and
 * hand the parent value on with a warp shuffle; the prepared series is staged
 * whole in shared memory.  Same source contract (fb_jit_slice, TH) and the same
 * semantics as fb_jit_slice_features.  FB_ENOSUP if the series does not fit
 * the shared-memory staging. */
typedef struct fb_jit_chain_geometry {
    int32_t blocks_per_series; /* warps working on the same series (trie blocks) */
    int32_t series_per_cta;
    int32_t rows_staged;       /* prepared dimensions staged per series */
    int32_t pad;               /* zero padding on both sides of a staged row (> max depth) */
} fb_jit_chain_geometry;
FB_API int fb_jit_chain_features(fb_jit_kernel *k, const fb_jit_chain_geometry *geo,
                                 const fb_batch *batch, const double *thr, int64_t n_thr,
                                 double *out, int64_t out_ld, int64_t col0, int sanitize,
                                 void *stream);

/* fruits/iss/semiring.py:103-125, :138-158: exp(+alpha g), exp(-alpha g) rows
 * of the exponential weighting for every distinct alpha:
 * out[r][2a + s][t] = exp((s ? -1 : 1) * alpha[a] * g[r][t]),  r < rows. */
FB_API int fb_exp_rows(const double *g, double *out, int64_t rows, int64_t t,
                       const float *alphas_h, int n_alphas, void *stream);

/* ISS.transform / batch_transform (fruits/iss/iss.py:118-185,
 * fruits/iss/semiring.py:93-201, :282-404): materialise every emitted
 * iterated sum of the plan as out[e][n][t] (the host compiles one plan per
 * chunk of words when the full tensor would not fit). */
FB_API int fb_iss_materialize(const fb_iss_plan *plan, const fb_batch *batch, double *out,
                              void *stream);

/* -- cosine weighted ISS (fruits/iss/cos.py) -- */

/* fruits/iss/cos.py:24-25: trig[f][0][t] = sin(pi t / (freq_f (t_len-1))),
 * trig[f][1][t] = cos(...); freqs is a DEVICE array of float32. */
FB_API int fb_cos_trig(const float *freqs, int n_freq, int64_t t, double *trig, void *stream);
/* Weight rows of the separable form of the cosine weighted ISS
 * (fruits/iss/cos.py:16-49 restated as a recurrence with exponent+1 states per
 * level, see fruits_b200/iss/cos.py): rows[r][t] = coeff * sin^a cos^b of
 * pi t / (freq_f (t_len-1)), spec = DEVICE int32 [n_rows][4] = (f, coeff, a, b),
 * freqs = DEVICE float32 [n_freq]. */
FB_API int fb_cos_rows(const float *freqs, int n_freq, int64_t t, const int32_t *spec,
                       int n_rows, double *rows, void *stream);
/* fruits/iss/cos.py:16-49, :289-333 in the separable form: all n_freq iterated
 * sums of ONE word, out[f][n][t].  word = DEVICE int32 [p][dw] exponents; rows
 * from fb_cos_rows; tab = DEVICE int32 [n_freq][ns + ns*ns + ns] row indices
 * (first level, inner levels [k'][k], last level / output), ns = exponent + 1.
 * FB_ENOSUP for words of more than 6 letters or exponents above 4 (the caller
 * then uses fb_coswiss_word). */
FB_API int fb_coswiss_sep_word(const double *X, int64_t n, int64_t d, int64_t t,
                               const int32_t *word, int p, int dw, const double *rows,
                               const int32_t *tab, int n_freq, int ns, int total, double *out,
                               void *stream);
/* _coswiss (fruits/iss/cos.py:16-49, :171-181) for one word: out[f][n][t].
 * word: device int32 [p][dw] exponent matrix (max_occ = the largest number of
 * occurrences in one letter, sum_d |word[k][d]|), weights: device int32
 * [n_terms][ncols] table of CosWISS._get_weightings (:265-287; ncols = 2p+1,
 * or 2p+3 with the total weighting; max_exp = its largest sin / cos exponent). */
FB_API int fb_coswiss_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word,
                           int p, int dw, int max_occ, int max_exp, const double *trig,
                           int n_freq, const int32_t *weights, int n_terms, int ncols,
                           double *out, void *stream);

/* fruits/iss/semiring.py:461-601 Bayesian semiring (max, times): iterated sums
 * of ONE word (exponents word[p][md], device memory; alpha[p] float32, device
 * memory) for all n series; the last `extended` prefixes are written to
 * out[extended][n][t] (semiring.py:478-483).  weight_mode FB_WEIGHT_*: g is
 * the weighting lookup (row stride g_ld, 0 = one row shared by all series).
 * A parallel running maximum over time -- exact, the maximum is associative. */
FB_API int fb_bayes_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word,
                         int p, int md, const float *alpha, const double *g, int64_t g_ld,
                         int weight_mode, int extended, double *out, void *stream);
/* The same block scan over T for one word of the Arctic (max, plus) semiring
 * (fruits/iss/semiring.py:282-338; the running maximum is exactly associative,
 * so the result is bit-identical to the time-serial kernels): used by
 * ISS.transform / fit for batches too small to fill the GPU with one lane per
 * trie node. */
FB_API int fb_arctic_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word,
                          int p, int md, const float *alpha, const double *g, int64_t g_ld,
                          int weight_mode, int extended, double *out, void *stream);

/* Arctic(argmax=True) for one word (fruits/iss/semiring.py:234-279
 * _arctic_argmax_single via Arctic._iterated_sum_fast :385-392): out holds
 * p + p(p+1)/2 rows [row][n][t] -- for every level k its running maximum
 * (row k + k(k+1)/2) followed by the k+1 position rows of the levels k..0 that
 * produced it.  alpha f32[p] (zeros when unweighted), g = lookup rows or NULL;
 * work = fb_arctic_argmax_workspace(n, t, p) bytes of device scratch. */
FB_API int64_t fb_arctic_argmax_workspace(int64_t n, int64_t t, int p);
FB_API int fb_arctic_argmax_word(const double *X, int64_t n, int64_t d, int64_t t,
                                 const int32_t *word, int p, int md, const float *alpha,
                                 const double *g, int64_t g_ld, double *out, void *work,
                                 void *stream);

/* -- preparateurs, lookups, raw-input cache -- */

/* fruits/cache.py:8-13 _increments on rows x t; pad_src != NULL keeps the
 * first k values of pad_src (INC(zero_padding=False), transform.py:72-75). */
FB_API int fb_increments(const double *X, const double *pad_src, double *out, int64_t rows,
                         int64_t t, int64_t k, void *stream);
/* fruits/preparation/transform.py:132-144 STD(separately=True):
 * stats[r] = (np.mean(row), np.std(row) + eps), numpy pairwise summation. */
FB_API int fb_row_stats(const double *X, double *stats, int64_t rows, int64_t t, int div_std,
                        double eps, void *stream);
FB_API int fb_standardize(const double *X, const double *stats, double *out, int64_t rows,
                          int64_t t, void *stream);
/* fruits/cache.py:25-40 _L1_sum / _L2_sum of dimension 0: out[n][t]. */
FB_API int fb_lsum(const double *X, double *out, int64_t n, int64_t d, int64_t t, int l2,
                   void *stream);
/* fruits/iss/weighting.py:151-158: optional r/(r[-1]+1e-5), NRM per row, * scale. */
FB_API int fb_nrm_scale(const double *in, double *out, int64_t rows, int64_t t, int relative,
                        double scale, void *stream);
/* fruits/cache.py:16-22 _coquantile: out[i] = count(S[i,:] <= q*S[i,-1]). */
FB_API int fb_coquantile(const double *S, int64_t *out, int64_t n, int64_t t, double q,
                         void *stream);

/* -- the remaining preparateurs (csrc/prep_more.cu): each writes a prepared
 *    copy [n][d'][t'] that the ISS kernels read like raw input -- */

/* DIL (fruits/preparation/filter.py:55-61), DOT (:185-190), PDD (:253-259),
 * CTS(pseudo_shift=True) (transform.py:940-941), WIN (filter.py:97-113):
 * out = X where keep[t] != 0 (keep == NULL: everywhere) and t lies in the
 * Python slice lo[i]+lo_off : hi[i] of its series (lo == hi == NULL:
 * everywhere), 0.0 elsewhere.  keep: uint8[t], lo / hi: int64[n] on the device. */
FB_API int fb_time_mask(const double *X, double *out, int64_t n, int64_t d, int64_t t,
                        const uint8_t *keep, const int64_t *lo, const int64_t *hi,
                        int64_t lo_off, void *stream);
/* CTS (transform.py:943-944): out[k] = x[min(k + shift, t - 1)] per row. */
FB_API int fb_time_shift(const double *X, double *out, int64_t rows, int64_t t, int64_t shift,
                         void *stream);
/* LAG (transform.py:291-298): rows x t -> (2 rows) x (2t - 1), lead row then lag row. */
FB_API int fb_lead_lag(const double *X, double *out, int64_t rows, int64_t t, void *stream);
/* MAV._backend (transform.py:233-239): out[k-1] = sum(x[k-width:k]) / width, 0 in front. */
FB_API int fb_moving_average(const double *X, double *out, int64_t rows, int64_t t,
                             int64_t width, void *stream);
/* RIN._backend (transform.py:447-468); kernel f64[d][width], ndim i32[n_out], dims i32[d] on
 * the device; pad = width reproduces adaptive_width (:537-543), else 0.  out[n][n_out][t]. */
FB_API int fb_random_increments(const double *X, const double *kernel, const int32_t *ndim,
                                const int32_t *dims, double *out, int64_t n, int64_t d,
                                int64_t t, int n_out, int width, int pad, void *stream);
/* JLD._backend (transform.py:651-670); kernel f64[sum(ndim)], bias f64[n_out]. */
FB_API int fb_dim_project(const double *X, const double *kernel, const double *bias,
                          const int32_t *ndim, const int32_t *dims, double *out, int64_t n,
                          int64_t d, int64_t t, int n_out, void *stream);
/* FFN._transform (transform.py:362-376); mean = fb_row_stats output (center) or NULL;
 * W1 f64[d_hidden][d], b1 f64[d_hidden], W2 f64[d_out][d_hidden].  out[n][d_out][t]. */
FB_API int fb_ffn(const double *X, const double *mean, const double *W1, const double *b1,
                  const double *W2, double *out, int64_t n, int64_t d, int64_t t, int d_hidden,
                  int d_out, int relu_out, void *stream);
/* RDW._transform (transform.py:601-602): x ** w[dim]. */
FB_API int fb_dim_pow(const double *X, const double *w, double *out, int64_t n, int64_t d,
                      int64_t t, void *stream);
/* RDW._fit (transform.py:592): out[j] = max_t(mean_i |x[i][j][t]|). */
FB_API int fb_abs_mean_max(const double *X, double *out, int64_t n, int64_t d, int64_t t,
                           void *stream);
/* RPE._backend (transform.py:859-875): X[n][2][t] rotated by k / den, den = T ** freq. */
FB_API int fb_rotate2(const double *X, double *out, int64_t n, int64_t t, double den,
                      void *stream);
/* SPE._transform (transform.py:790-802): the argument of the wave (src == NULL: k / den in
 * one row; else src / den, or src / src[r][t-1] ** freq if per_row_last), through np.sin
 * if apply_sin.  den = T ** freq from the host. */
FB_API int fb_spe_range(const double *src, double *out, int64_t rows, int64_t t, double den,
                        double freq, int per_row_last, int apply_sin, void *stream);
/* SPE._transform (transform.py:803-810): X[x_rows][d][t] * (or +) wave[wave_rows][1][t] with
 * numpy's broadcasting (equal row counts, or one of them 1); out has the larger row count. */
FB_API int fb_wave_embed(const double *X, const double *wave, double *out, int64_t x_rows,
                         int64_t wave_rows, int64_t d, int64_t t, int additive, void *stream);
/* QTC._transform (transform.py:990-1001): np.where(X > q, bound, X) (lower: X < q). */
FB_API int fb_clip_where(const double *X, double *out, int64_t total, double q, double bound,
                         int lower, void *stream);

/* -- sieves on materialised arrays (stand-alone seeds, general cuts, fit) -- */

/* fruits/sieving/increment.py:63-71 _pre_transform: inc > 0 increments,
 * inc < 0 cumulative sums, inc == 0 copy. */
FB_API int fb_pretransform(const double *Y, double *out, int64_t rows, int64_t t, int inc,
                           void *stream);
#define FB_SIEVE_NPI 0
#define FB_SIEVE_MPI 1
#define FB_SIEVE_MAX 2
#define FB_SIEVE_MIN 3
#define FB_SIEVE_XPI 4
#define FB_SIEVE_LPI 5
#define FB_SIEVE_END 6
#define FB_SIEVE_CUR 7   /* segment.py:228-274 (also what AVG :277-317 and STD :320-358 run) */
/* fruits/sieving/segment.py:107-225, increment.py:101-239 backends on
 * V[rows][ld]; cuts[rows][nc] sorted with first column 0 (NULL: {0, t});
 * q[nq] sorted thresholds; out[r*out_ld + col0 + seg*(nq-1) + k]. */
FB_API int fb_segment_sieve(const double *V, int64_t ld, const int64_t *cuts, int nc,
                            const double *q, int nq, int kind, double *out, int64_t out_ld,
                            int64_t col0, int64_t rows, int64_t t, void *stream);
/* fruits/sieving/implicit.py:114-129 PPV._transform (segments = 0 / 1);
 * segments | 2: :169-190 CPV._transform (connected components). */
FB_API int fb_ppv(const double *V, int64_t ld, const double *q, int nq, int segments, double *out,
                  int64_t out_ld, int64_t col0, int64_t rows, int64_t t, void *stream);
/* np.nan_to_num(a, nan=0.0) in place (fruits/fruit.py:172). */
FB_API int fb_nan_to_num(double *a, int64_t total, void *stream);

/* Multi-GPU assembly of the feature matrix (north_star (3)): copy
 * a rows x cols block of features into the same block of an NVSwitch multicast
 * mapping (NVLS) with multimem.st, i.e. into the matrix of every rank at once.
 * Used for the slices that are not written by a generated kernel (those store
 * through the mapping themselves, flag bit 1 of `sanitize` in fb_jit_*_features). */
FB_API int fb_multimem_copy(const double *src, int64_t src_ld, double *mc_dst, int64_t dst_ld,
                            int64_t rows, int64_t cols, void *stream);

/* -- fit: exact order statistics for np.quantile (segment.py:66-75) -- */
FB_API int64_t fb_order_stats_workspace(int64_t P);
/* P problems of M doubles (problem p at V + p*ldp): lo[p] = x_(k),
 * hi[p] = x_(min(k+1, M-1)) of the ascending order; NaN if any NaN. */
FB_API int fb_order_stats(const double *V, int64_t ldp, int64_t P, int64_t M, int64_t k,
                          double *lo, double *hi, void *work, void *stream);

/* The same for n_sel (<= 4) selections per problem in three reads of the
 * data: selection s ranks the inc[s]-fold zero-padded increments (0..2; 0 =
 * the values) of the rows of length t (IncrementSieve._pre_transform,
 * increment.py:63-71, formed on the fly) and returns x_(k[s]), x_(k[s]+1) in
 * lo/hi[p*n_sel + s].  done[p*n_sel + s] = 0: more equal values than the
 * candidate list holds, repeat that selection with fb_order_stats. */
FB_API int64_t fb_order_stats_multi_workspace(int64_t P, int n_sel);
FB_API int fb_order_stats_multi(const double *V, int64_t ldp, int64_t P, int64_t M, int64_t t,
                                int n_sel, const int32_t *inc, const int64_t *k, double *lo,
                                double *hi, int32_t *done, void *work, void *stream);

/* -- row-sharded fit: the thresholds are quantiles over the whole
 * fit sample (fruits/sieving/segment.py:66-75) while every rank holds a share of
 * its rows.  The selections run in phases over the LOCAL rows; between the phases
 * the host sums (MINs) small regions of the workspace over the ranks, so all
 * ranks walk the same radix buckets and end with the same order statistics:
 *   fb_order_stats_dist   three reads of the local data (12 + 12 bits) and five
 *                         8-bit passes over the local candidate lists, phases 0..9;
 *                         layout = byte offsets of {hist u32[P*n_sel*4096] (SUM after
 *                         phases 0, 1), h256 u32[P*n_sel*256] (SUM after 3..7), sums
 *                         i64[P*n_sel*2] (SUM after 2, 8), mins i64[P*n_sel*4] (MIN after
 *                         2, 8)}, workspace bytes
 *   fb_order_stats_dist8  eight 8-bit passes over materialised local values, phases
 *                         0..9; layout = {hist u32[P*256] (SUM after 0..7), sums
 *                         i64[P*2] (SUM after 8), mins i64[P] (MIN after 8)}, bytes.
 * k is the rank among the m_global values of all ranks (< 2^31). */
FB_API int fb_order_stats_dist_layout(int64_t P, int n_sel, int64_t *layout);
FB_API int fb_order_stats_dist(int phase, const double *V, int64_t ldp, int64_t P,
                               int64_t m_local, int64_t m_global, int64_t t, int n_sel,
                               const int32_t *inc, const int64_t *k, double *lo, double *hi,
                               int32_t *done, void *work, void *stream);
FB_API int fb_order_stats_dist8_layout(int64_t P, int64_t *layout);
FB_API int fb_order_stats_dist8(int phase, const double *V, int64_t ldp, int64_t P,
                                int64_t m_local, int64_t m_global, int64_t k, double *lo,
                                double *hi, void *work, void *stream);

/* -- measurement -- */
/* fp64 FMA microbenchmark (grid x 256 threads x iters*64 DFMA each; out holds
 * grid*256 doubles): the measured fp64 roof bench.py reports against. */
FB_API int fb_fp64_peak(double *out, int grid, int iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FRUITS_B200_H */
