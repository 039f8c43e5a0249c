/*
 * ocm_simca.h -- C ABI of the B200-native SIMCA / DD-SIMCA hot path (libocm_simca_b200.so).
 *
 * The reference (TEAM-AIOLY/OCM-VAE-SIMCA) is pure Python and has no FFI of its own
 * (SURVEY.md §8(b)); its boundary for this path is the arithmetic that utils/SIMCA.py
 * delegates to numpy / scikit-learn.  Each entry point below replaces one such call
 * site (cited per function) and is what a ctypes / cffi / pybind stub on the reference
 * side binds -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t on a CUDA failure or a
 *     negative OCM_E* code on bad arguments; ocm_last_error() gives a thread-local message;
 *   - all array pointers are DEVICE pointers unless the parameter name starts with h_;
 *   - matrices are row-major with an explicit leading dimension in ELEMENTS;
 *   - the caller owns every buffer; nothing is allocated, freed or synchronised inside
 *     (exception: the *_host convenience calls at the end, which say so);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - calls on distinct streams / buffers are thread-safe.
 */
#ifndef OCM_SIMCA_H_
#define OCM_SIMCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCM_ABI_VERSION 2   /* 2: ocm_score_params.x_rms */

#define OCM_EINVAL (-1)   /* bad argument (NULL pointer, non-positive size, unknown enum) */
#define OCM_ELIMIT (-2)   /* size outside what the kernels support (e.g. n_components > 64) */
#define OCM_EALIGN (-3)   /* pointer / leading dimension alignment not supported by this path */
#define OCM_ETIMEOUT (-4) /* a tensor-core kernel of an EARLIER launch abandoned a bounded mbarrier wait.  Scoring kernels
                             trap (that launch fails as a whole, CUDA reports it at the next synchronisation, every later
                             tensor-core call returns this code); the Gram kernel writes NaN tiles and the code is
                             reported once.  The call that returns it has launched nothing. */

/* preprocessing modes */
#define OCM_PRE_SNV       1   /* (x - mean_row) / (std_row(ddof=0) + eps)      simca_nuts.py:48-50 */
#define OCM_PRE_CENTER    2   /* x - mu_col                                     sklearn PCA centring, SIMCA.py:64,:127 */
#define OCM_PRE_AUTOSCALE 3   /* (x - mu_col) / sigma_col                       vae_model.py:124 */
#define OCM_PRE_SNV_CENTER 4  /* SNV then column centring in one pass           */

/* decision rules (utils/SIMCA.py:131-145) */
#define OCM_RULE_SIM 0   /* max(a, b)        */
#define OCM_RULE_ALT 1   /* sqrt(a^2 + b^2)  */
#define OCM_RULE_CI  2   /* a + b            */
#define OCM_RULE_DD  3   /* a + b with a = t2dof*T2/h0, b = qdof*Q/v0 (same arithmetic as CI) */

int         ocm_abi_version(void);
const char* ocm_last_error(void);
/* number of kernel launches issued through this library by the calling process (bench.py's gpu_launches) */
uint64_t    ocm_launch_count(void);

/* ---- a1 / a2: streaming preprocessing ------------------------------------------------
 * Y = pre(X).  mu / sigma are column vectors [p] (may be NULL where the mode does not use
 * them).  In-place (Y == X, ldy == ldx) is allowed. */
int ocm_preprocess_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t mode,
                       const float* mu, const float* sigma, float eps,
                       float* Y, int64_t ldy, void* stream);

/* ---- f1: Savitzky-Golay filter, optionally after row SNV (simca_nuts.py:48-52, utils/data_utils.py:57-59:
 * scipy.signal.savgol_filter(window_length=5, polyorder=2, deriv=1, mode='interp')).  h_coef [window] (HOST) are the
 * interior taps g, y[i] = sum_d g[d] x[i + d - window/2]; edge_l / edge_r [window/2][window] (device) map the first /
 * last `window` samples to the window/2 outputs at each end (the polynomial-fit edges of mode 'interp').
 * Out of place only. */
int ocm_savgol_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t snv, float eps,
                   const float* h_coef, int32_t window, const float* edge_l, const float* edge_r,
                   float* Y, int64_t ldy, void* stream);

/* ---- a2 / a4: column moments of selected rows -----------------------------------------
 * colsum[j] += sum_r (x[r][j] - shift[j]),  colsumsq[j] += sum_r (x[r][j] - shift[j])^2 in fp64.
 * row_idx (int64, may be NULL = rows 0..n-1) gathers rows; shift / colsumsq may be NULL.
 * The outputs are ACCUMULATED into (zero them first). */
int ocm_colstats_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                     const double* shift, double* colsum, double* colsumsq, void* stream);
int ocm_colstats_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                     const double* shift, double* colsum, double* colsumsq, void* stream);

/* One pass for the bulk fit: colsum[j] += sum_r x[r][j] (fp64), colmin[j] = min(colmin[j], min_r x[r][j]),
 * colmax[j] likewise (fp32, rounded outwards for fp64 rows).  The caller initialises colmin to +inf and colmax
 * to -inf; either may be NULL. */
int ocm_colstats_minmax_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                            double* colsum, float* colmin, float* colmax, void* stream);
int ocm_colstats_minmax_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                            double* colsum, float* colmin, float* colmax, void* stream);

/* ---- a4: partitioned Gram accumulation (replaces the SVD inside sklearn PCA.fit,
 * utils/SIMCA.py:64-66,:75) ------------------------------------------------------------
 * G[i][j] += sum_r (x[r][i]-shift[i]) (x[r][j]-shift[j]) for j >= i (UPPER triangle only),
 * fp64 accumulation on the FP64 tensor pipe (DMMA; FP64 CUDA cores for tiny problems).
 * G is p x p row-major with leading dimension p. */
int ocm_gram_f64acc_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                        const double* shift, double* G, void* stream);
int ocm_gram_f64acc_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                        const double* shift, double* G, void* stream);

/* ---- a4, bulk: the same Gram on the 5th-generation tensor cores, exact integer arithmetic (csrc/gram_tc.cu).
 * Rows are quantised to 31-bit fixed point per column, q = rint((x - shift) 2^sc) 2^-sc with 2^sc chosen from
 * colmin / colmax (|q 2^sc| <= 2^30), split into four balanced int8 digit planes and contracted with
 * tcgen05.mma.kind::i8 (int32 accumulators in TMEM, chains <= 32 704 rows, flushed into fp64):
 *   G[i][j]   += sum_r q[r][i] q[r][j]   (j >= i, UPPER triangle only)
 *   colsum[j] += sum_r q[r][j]           (exact: int64 accumulation; may be NULL)
 * products = 16: every digit product -> the exact Gram of the quantised rows; 13: products of weight < 256^2
 * dropped (<= 2^-37 max_i max_j per row); 10: weight < 256^3 dropped (<= 2^-28).
 * colmin / colmax [p] fp32 bound the column values (ocm_colstats_minmax_*); flags [2] int64 (device, may be NULL)
 * receives {values clamped because they exceeded the bounds, non-finite inputs} -- both must be 0 for the result to
 * hold.  work: device scratch of ocm_gram_tc_workspace_bytes(n, p) bytes, 1024-byte aligned. */
int ocm_gram_tc_workspace_bytes(int64_t n, int32_t p, int64_t* bytes);
int ocm_gram_tc_f32(const float* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                    const double* shift, const float* colmin, const float* colmax, int32_t products,
                    void* work, int64_t work_bytes, double* G, double* colsum, int64_t* flags, void* stream);
int ocm_gram_tc_f64(const double* X, int64_t ldx, const int64_t* row_idx, int64_t n, int32_t p,
                    const double* shift, const float* colmin, const float* colmax, int32_t products,
                    void* work, int64_t work_bytes, double* G, double* colsum, int64_t* flags, void* stream);

/* Combine per-partition moments into a covariance:  with weights w[g] (0 / 1 to select partitions, -1 to DOWNDATE a
 * held-out part from a pool: CV folds)
 *   N = sum w_g n_g,  s = sum w_g colsum_g,  G = sum w_g G_g,  d = s/N,
 *   mean = shift + d,  C = (G - N d d^T)/(N-1)  (full symmetric p x p written to C);
 *   delta[p] (required scratch/output) receives d, mean[p] may be NULL.
 * Gparts [ngroups][p][p] (upper triangles valid), sparts [ngroups][p], h_counts / h_weights
 * are HOST arrays of length ngroups.  Used for plain fits (ngroups = 1), the multi-GPU
 * all-reduced fit and CV folds (weight 0 on the held-out fold). */
int ocm_cov_from_parts_f64(const double* Gparts, const double* sparts, const int64_t* h_counts,
                           const double* h_weights, int32_t ngroups, int32_t p,
                           const double* shift, double* C, double* mean, double* delta,
                           void* stream);

/* ---- a4 / a12: batched symmetric eigensolver for small matrices (m <= 64), parallel cyclic Jacobi, one CTA per
 * matrix.  H [batch][m][m] fp64 (symmetrised internally), evals [batch][m] DESCENDING, evecs [batch][m][m] with
 * eigenvectors as COLUMNS in the same order.  It is the Rayleigh-Ritz step of the batched subspace iteration that
 * fits all CV folds at once (replaces the per-(fold, LV) SVDs of utils/CVSIMCA.py:183-184). */
int ocm_eigh_small_f64(const double* H, int32_t m, int32_t batch, double* evals, double* evecs, void* stream);

/* ---- a4 / a12: steps of the batched block subspace iteration that yields the k leading eigenpairs of the covariance
 * (the loadings: Vt of the SVD at utils/SIMCA.py:64,:75) and the eigenvalue tail sums of utils/SIMCA.py:189-205.
 * Row-major fp64; C [batch][p][p] symmetric; blocks [batch][p][kb], kb <= 64 (csrc/eig_sub.cu).
 *   symm_mul:       Z = C Q                                   (FP64 tensor pipe)
 *   tn_small:       H = A^T B, [batch][ka][kb]
 *   chol_solve:     L = chol(sym(G)) lower, Y <- Y L^-T in place; info[b] = 0 or the 1-based index of a non-positive pivot
 *   mul_small:      Y = Z U, U [batch][kb][ku]
 *   ritz_residual:  res2[b][c] = sum_rows (ZU[row][c] - X[row][c] w[b][c])^2
 *   tail_sums:      R = C - sum_{a<k} w_a x_a x_a^T (written to R), R2 = upper triangle of R^2 (scratch),
 *                   tails[b] = {tr R, |R|_F^2, tr R^3} */
int ocm_eig_symm_mul_f64(const double* C, const double* Q, double* Z, int32_t p, int32_t kb, int32_t batch, void* stream);
int ocm_eig_tn_small_f64(const double* A, const double* B, double* H, int32_t p, int32_t ka, int32_t kb,
                         int32_t batch, void* stream);
int ocm_eig_chol_solve_f64(const double* G, double* Y, double* L, int32_t* info, int32_t p, int32_t kb,
                           int32_t batch, void* stream);
int ocm_eig_mul_small_f64(const double* Z, const double* U, double* Y, int32_t p, int32_t kb, int32_t ku,
                          int32_t batch, void* stream);
int ocm_eig_ritz_residual_f64(const double* ZU, const double* X, const double* w, double* res2, int32_t p,
                              int32_t kb, int32_t batch, void* stream);
int ocm_eig_tail_sums_f64(const double* C, const double* X, const double* w, int32_t p, int32_t kb, int32_t k,
                          int32_t batch, double* R, double* R2, double* tails, void* stream);

/* ---- a4: ALL eigenvalues of a symmetric fp64 matrix (the `eigs_all` array of utils/SIMCA.py:66,:89, computed lazily):
 * Householder tridiagonalisation + Sturm-sequence bisection (csrc/eig_full.cu).  A [p][p] is OVERWRITTEN; work holds
 * at least 2 p + 1 doubles; d, e [p] receive the tridiagonal, w [p] the eigenvalues in DESCENDING order. */
int ocm_eigvals_sym_f64(double* A, int32_t p, double* work, double* d, double* e, double* w, void* stream);

/* ---- a5 / a9 / a10: fused scoring (replaces pca.transform + inverse_transform + the
 * two reductions of SIMCA.predict / SIMCA.transform, utils/SIMCA.py:101-145) -----------
 * For each row x:  [optional SNV] -> xc = x - mu -> t = P xc -> r = xc - P^T t ->
 *   T2 = t^T Sinv t,  Q = |r|^2,  a = T2*ct, b = Q*cq, d = rule(a,b), accept = d < dlim.
 * P is [A][p] row-major (orthonormal rows), Sinv is [A][A] full symmetric.
 * T2 / Q / accept / scores may each be NULL; scores [n][A] (the PCA scores t, exported only
 * for the reference's `_model[cls]["T"]`) is off on the bulk path.  Only distances and
 * decisions go back to HBM. */
typedef struct {
    int32_t A;          /* number of components (1..64) */
    int32_t rule;       /* OCM_RULE_* */
    int32_t snv;        /* 1: apply row SNV (eps below) before centring */
    float   snv_eps;
    float   ct, cq;     /* reduced-distance scales (1/T2_limit, 1/Q_limit; dof/scale for dd) */
    float   dlim;       /* critical distance (strict <) */
    float   x_rms;      /* 0 = unknown.  > 0: root-mean-square size of one centred element, sqrt(trace(cov) / p) of the
                           data the model was fitted on, AND a promise that |P_ij| <= 1 (unit-norm loadings).  The
                           tcgen05 kernel then takes its fp16 operand mode (power-of-two prescales chosen from x_rms,
                           2-term split of the spectra): ~10 % faster, same 1e-5 parity; rows whose operands leave the
                           fp16 range (|x - mu| > ~4000 x_rms) come out non-finite and are recomputed in fp32 by a
                           fix-up kernel on the same stream.  With 0 the exact-at-any-scale bf16 x 3 mode runs. */
} ocm_score_params;

int ocm_simca_score_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                        const float* mu, const float* P, const float* Sinv,
                        const ocm_score_params* h_params,
                        float* T2, float* Q, uint8_t* accept, float* scores, void* stream);
/* fp64 rows (numpy's default dtype).  Bulk shapes (the fp32 gate below on n, p, n_components; 16-byte aligned rows
 * and mean; scores == NULL) are centred -- and SNV-normalised when asked -- in fp64, rounded once to fp32 into a bounded
 * stream-ordered scratch buffer (1 GB chunks by default, a memory pool of the library's own) and scored chunk by chunk on
 * the tcgen05 kernels; anything
 * else runs on the generic kernel, which rounds at the same point. */
int ocm_simca_score_f64in(const double* X, int64_t n, int32_t p, int64_t ldx,
                          const double* mu, const float* P, const float* Sinv,
                          const ocm_score_params* h_params,
                          float* T2, float* Q, uint8_t* accept, float* scores, void* stream);
int ocm_simca_score_f64in_bulk_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, const void* X, const void* mu);
/* fp32 rows with fused SNV at shapes whose plain form the tcgen05 kernels take but whose fused form they do not
 * (1152 < p <= 2048): ocm_simca_score_f32 stages SNV + centring (ocm_preprocess_f32, OCM_PRE_SNV_CENTER) through the same
 * scratch and scores the chunk on the cluster kernel.  The scratch pool is released by ocm_simca_host_release. */
int ocm_simca_score_snv_staged_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, const void* X, const void* mu);
int ocm_simca_score_snv_staged_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                                   const float* mu, const float* P, const float* Sinv,
                                   const ocm_score_params* h_params,
                                   float* T2, float* Q, uint8_t* accept, void* stream);
/* Raw spectra -> distances and decisions in one call: [row SNV +] Savitzky-Golay filter (the arguments of ocm_savgol_f32)
 * into a bounded scratch chunk, then the bulk scoring kernels against a model fitted in the filtered space
 * (simca_nuts.py:48-52 followed by utils/SIMCA.py:120-154, without the second full-size matrix).  h_params->snv must be 0. */
int ocm_simca_score_savgol_f32(const float* X, int64_t n, int32_t p, int64_t ldx, int32_t snv, float snv_eps,
                               const float* h_coef, int32_t window, const float* edge_l, const float* edge_r,
                               const float* mu, const float* P, const float* Sinv, const ocm_score_params* h_params,
                               float* T2, float* Q, uint8_t* accept, void* stream);
/* the fp64 contract above, forced onto the generic kernel */
int ocm_simca_score_simt_f64in(const double* X, int64_t n, int32_t p, int64_t ldx,
                               const double* mu, const float* P, const float* Sinv,
                               const ocm_score_params* h_params,
                               float* T2, float* Q, uint8_t* accept, float* scores, void* stream);
/* same contract, forced onto the FP32-pipe kernel (any shape; the dispatcher above picks the
 * tcgen05 kernel when the shape allows it) */
int ocm_simca_score_simt_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                             const float* mu, const float* P, const float* Sinv,
                             const ocm_score_params* h_params,
                             float* T2, float* Q, uint8_t* accept, float* scores, void* stream);

/* the register-resident "row team" FP32-pipe kernel (bulk shapes: p in {512,1024,2048}, even
 * n_components <= 12, 16-byte aligned rows, no fused SNV) and its shape gate */
int ocm_simca_score_rowteam_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
int ocm_simca_score_rowteam_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                                const float* mu, const float* P, const float* Sinv,
                                const ocm_score_params* h_params,
                                float* T2, float* Q, uint8_t* accept, void* stream);

/* 1 when ocm_simca_score_f32 would take a tcgen05/TMA kernel for this shape (bulk path: fp32 rows, 64 <= p <= 2048, p a
 * multiple of 4, n_components <= 16, 16-byte aligned rows, n >= 32768; fused SNV up to p = 1152), else 0. */
int ocm_simca_score_tc_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
/* the tcgen05 kernels themselves: p <= 1152 runs one CTA per 128-spectrum tile (score_tc.cu; fp16 operand mode when
 * h_params->x_rms > 0), wider spectra the cluster kernel below with four CTAs per tile */
int ocm_simca_score_tc_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                           const float* mu, const float* P, const float* Sinv,
                           const ocm_score_params* h_params,
                           float* T2, float* Q, uint8_t* accept, void* stream);
/* cluster kernel (score_tc2.cu): the wavelengths of a tile are split over 1 / 2 / 4 CTAs of a thread-block cluster
 * (<= 512 each, p <= 2048), both passes run concurrently, partial scores meet through distributed shared memory.
 * inject_polls: 0, or the bounded-wait limit of a launch whose first CTA drops one TMA stage (test hook). */
int ocm_simca_score_tc2_supported(int64_t n, int32_t p, int64_t ldx, int32_t A, int32_t snv, const void* X);
int ocm_simca_score_tc2_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                            const float* mu, const float* P, const float* Sinv,
                            const ocm_score_params* h_params,
                            float* T2, float* Q, uint8_t* accept, int32_t inject_polls, void* stream);
/* bring-up aid: copies the 16 words of the mapped protocol status block (word 0 = status, the rest a role trace when
 * the library was built with OCM_TC2_TRACE) */
int ocm_simca_score_tc_debug_words(uint32_t* out16);
/* Synchronises the device, then returns the protocol status of the tcgen05 kernels (scoring and Gram):
 * 0 = every launch so far completed its hand-offs, > 0 = a bounded mbarrier wait was abandoned (1, 2: a scoring kernel
 * trapped, the launch failed and the context is lost; >= 10: the Gram kernel finished with NaN tiles, cleared by this
 * call), -1 = CUDA error without a status code.  The scoring / Gram entry points check the same word (without
 * synchronising) before they launch and return OCM_ETIMEOUT. */
int ocm_simca_score_tc_status(void);
/* test hook: the next tcgen05 scoring launch drops one TMA stage in its first CTA, so that the bounded wait gives up
 * after `max_polls` polls (>= 16).  Exists so that the failure path is exercised by the test-suite. */
int ocm_simca_score_tc_inject_stall(int32_t max_polls);

/* ---- a12: all nested models A = 1..Amax in one pass (CV grid, CVSIMCA.py:157-200) ------
 * inv_eig[Amax] = 1/lambda_a (diagonal Sinv of a PCA model).  Outputs are [n][Amax]:
 *   T2[i][A-1] = sum_{a<=A} t_a^2 inv_eig_a,  Q[i][A-1] = |r_Amax|^2 + sum_{a>A} t_a^2.
 * ct / cq / dlim are per-A arrays [Amax] (device); accept may be NULL. */
int ocm_simca_score_multiA_f32(const float* X, int64_t n, int32_t p, int64_t ldx,
                               const float* mu, const float* P, const float* inv_eig,
                               int32_t Amax, int32_t rule, int32_t snv, float snv_eps,
                               const float* ct, const float* cq, const float* dlim,
                               float* T2, float* Q, uint8_t* accept, void* stream);
int ocm_simca_score_multiA_f64in(const double* X, int64_t n, int32_t p, int64_t ldx,
                                 const double* mu, const float* P, const float* inv_eig,
                                 int32_t Amax, int32_t rule, int32_t snv, float snv_eps,
                                 const float* ct, const float* cq, const float* dlim,
                                 float* T2, float* Q, uint8_t* accept, void* stream);

/* ---- a11 / f3: conformity counts on device (utils/SIMCA.py:238-266) --------------------
 * counts[m][4] += {TP, TN, FP, FN} for each of the m decision columns of accept[n][m],
 * is_target[n] = 1 where y_true == class_index. */
int ocm_conformity_counts(const uint8_t* accept, const uint8_t* is_target, int64_t n, int32_t m,
                          int64_t* counts, void* stream);

/* ---- a14: VAE-SIMCA latent / reconstruction scoring (VAE_SIMCA.py:359-365,
 * utils/final_vaesimca.py:479-492, vae_model.py:164) ------------------------------------
 *   d2[i] = (z_i - m)^T Sinv (z_i - m)     z [n][d] fp32, Sinv [d][d] full symmetric
 *   q[i]  = sum_j (x[i][j] - xrec[i][j])^2 x, xrec [n][p] (xrec fp32 or bf16)
 * x / xrec may be NULL (then q is not written); z may be NULL (then d2 is not written). */
int ocm_latent_score_f32(const float* z, int64_t n, int32_t d, int64_t ldz,
                         const float* m, const float* Sinv,
                         const float* x, const void* xrec, int32_t xrec_is_bf16,
                         int32_t p, int64_t ldx, int64_t ldxrec,
                         float* d2, float* q, void* stream);

/* ---- a15 / config 5: the elementwise-bound pieces of the ConvVAE1D scoring pass (vae_model.py:6-129 driven by
 * utils/final_vaesimca.py:471-533), for channels-last bf16 activations [n][L][C].  act: 0 none, 1 ELU, 2 GELU (erf).
 * ocm_vae_conv_in_bf16: first encoder block in one pass -- (x - spec_mean) / spec_std, Conv1d(1 -> C, K taps, stride 1,
 * zero padding K/2; w [C][K], bias [C], eval-mode BatchNorm folded in by the caller), activation, bf16 output
 * [n][L][C].  C in {16, 32, 64}, K odd <= 15.
 * ocm_bias_act_bf16: Y[r][c] <- act(Y[r][c] + bias[c]) in place on a dense bf16 [rows][C] array, C a multiple of 8
 * (the epilogue cuDNN does not fuse for ELU / GELU). */
int ocm_vae_conv_in_bf16(const float* X, int64_t n, int32_t L, int64_t ldx, const float* spec_mean, const float* spec_std,
                         const float* w, const float* bias, int32_t C, int32_t K, int32_t act, void* out_bf16, void* stream);
int ocm_bias_act_bf16(void* Y_bf16, int64_t rows, int32_t C, const float* bias, int32_t act, void* stream);
/* Tail of the decoder in one pass (vae_model.py:95-129 + the reconstruction Q of final_vaesimca.py:484-492): Y [n][Ly][C]
 * bf16 = raw output of the last transposed convolution; a = act(Y + bias), s = sum_c a_c w_out[c] + b_out (the 1x1 output
 * convolution), x_rec_l = s spec_std_l + spec_mean_l (l >= Ly: s = 0), q = sum_{l < L} (X_l - x_rec_l)^2.  xrec_bf16
 * [n][L] may be NULL.  C in {16, 32, 64}. */
int ocm_vae_tail_q_bf16(const void* Y_bf16, int64_t n, int32_t Ly, int32_t C, const float* bias, int32_t act,
                        const float* w_out, float b_out, const float* X, int64_t ldx, int32_t L,
                        const float* spec_mean, const float* spec_std, float* q, void* xrec_bf16, void* stream);

/* ---- host-buffer convenience (the end-to-end call bench.py times as `e2e`) -------------
 * Scores h_X (HOST, pageable or pinned) in row chunks (chunk_rows <= 0: 65536), overlapping H2D copies with the
 * kernel on two internal streams, and writes h_T2 / h_Q / h_accept (HOST).  Synchronises before returning.  Its
 * device staging buffers and streams are a per-device workspace kept between calls (calls on one device are serialised);
 * ocm_simca_host_release() frees it.
 * Pinned fp32 rows are copied straight from the caller's buffer.  Pageable rows (a numpy array) and fp64 rows go through
 * two pinned bounce buffers that host threads fill in parallel (OCM_HOST_THREADS, default min(cores, 16)) while the
 * previous chunk is on the link; fp64 rows are centred in fp64 and rounded once during that fill -- (float)(x - mu), the
 * rounding point of ocm_simca_score_f64in -- so 4 bytes per element cross the link (with fused SNV: rounded, then
 * normalised and centred by the kernel).  This is what utils/SIMCA.py:120-154 (predict) and :101-117 (transform) amount to
 * for a numpy matrix that does not fit, or is not worth holding, in device memory. */
int ocm_simca_score_host_f32(const float* h_X, int64_t n, int32_t p, int64_t ldx,
                             const float* h_mu, const float* h_P, const float* h_Sinv,
                             const ocm_score_params* h_params, int64_t chunk_rows,
                             float* h_T2, float* h_Q, uint8_t* h_accept);
int ocm_simca_score_host_f64in(const double* h_X, int64_t n, int32_t p, int64_t ldx,
                               const double* h_mu, const float* h_P, const float* h_Sinv,
                               const ocm_score_params* h_params, int64_t chunk_rows,
                               float* h_T2, float* h_Q, uint8_t* h_accept);
/* Pageable (or pinned) host rows -> device rows through the same workspace: n rows of row_bytes bytes, strides in bytes.
 * Pageable sources are copied into the pinned bounce buffers by the host threads while the previous chunk is on the link
 * (a plain cudaMemcpy of pageable memory is staged by the driver on one thread: ~11 GB/s on the B200 box against ~37).
 * Blocking: the data is complete on return. */
int ocm_upload_rows(const void* h_src, int64_t n, int64_t row_bytes, int64_t src_stride_bytes,
                    void* d_dst, int64_t dst_stride_bytes);
int ocm_simca_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* OCM_SIMCA_H_ */
