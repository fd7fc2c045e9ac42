/*
 * mapstorch_b200 -- C ABI of the B200-native per-pixel XRF map-fitting hot path.
 *
 * The reference (xyin-anl/MapsTorch) has no FFI: its boundary is plain Python
 * (SURVEY.md section 8b).  These entry points are what a `mapstorch`-compatible
 * Python package binds with ctypes; each one names the
 * reference code it replaces.  Conventions:
 *   - every pointer suffixed _dev is a DEVICE pointer owned by the caller (torch
 *     tensors); pointers suffixed _host are host memory; the library allocates
 *     nothing persistent;
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered and
 *     asynchronous, the caller synchronises;
 *   - return value: MT_OK or a negative MT_ERR_* code; mt_last_error() gives the
 *     text of the last failure on the calling thread.  No exceptions cross the ABI;
 *   - there is NO CPU fallback: every compute entry point fails with
 *     MT_ERR_NO_DEVICE unless the current device is compute capability 10.x.
 */
#ifndef MAPSTORCH_B200_H
#define MAPSTORCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_ABI_VERSION 1

#define MT_OK 0
#define MT_ERR_INVALID (-1)     /* bad argument (null pointer, size, alignment) */
#define MT_ERR_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define MT_ERR_UNSUPPORTED (-3) /* shape or dtype outside what the kernels cover */
#define MT_ERR_NO_DEVICE (-4)   /* current device is not sm_100 */

/* layout of the amplitude array x */
#define MT_X_ELEMENT_MAJOR 0 /* x[e * ld_x + p]: one contiguous map per component */
#define MT_X_PIXEL_MAJOR 1   /* x[p * ld_x + e]: one contiguous 16-byte aligned record per pixel; ld_x % 4 == 0,
                                ld_x >= e_pad (columns >= n_comp are written as zeros) */

/* storage type of a spectral volume */
#define MT_DTYPE_F32 0
#define MT_DTYPE_F16 1

/* ---- model terms (kernel K1) ------------------------------------------------
 * One additive term of the MAPS forward model, evaluated per energy channel in
 * the reference's fp32 operation order (map.py:57-69).  `delta = sign*(ev - e)`.
 *   MT_TERM_GAUSS: amp * (p1 * exp(-0.5 * (delta/p0)^2))          p0=sigma p1=gain/(sigma*sqrt(2pi))
 *   MT_TERM_STEP : amp * (p1 * erfc(delta/p0))                    p0=sqrt2*sigma p1=gain/2/E
 *   MT_TERM_TAIL : amp * (p3 * (delta<0 ? exp(delta/p2)*v : v)),  v=erfc(delta/p0 + p1)
 *                  p0=sqrt2*sigma p1=1/(gamma*sqrt2) p2=gamma*sigma p3=gain/2/gamma/sigma/exp(-0.5/gamma^2)
 * Terms of one output column are stored contiguously and are accumulated in order.
 */
#define MT_TERM_GAUSS 0
#define MT_TERM_STEP 1
#define MT_TERM_TAIL 2

typedef struct MtTerm {
    int32_t kind;  /* MT_TERM_* */
    float sign;    /* +1 or -1 (map.py:121 passes -delta_energy for the high-energy Compton tail) */
    float e;       /* line energy, keV, already rounded to fp32 */
    float p0, p1, p2, p3;
    float amp;     /* faktor of the term (ratio * 10**logA / normalisation [* f_step | f_tail]) */
} MtTerm;

/* Extra destinations of the element maps on OTHER GPUs of the box (NVLink peer memory): every value
 * the kernels store to x_dev[i] is also stored to peer[k][i] for k < n_peers and, if multicast is
 * non-NULL, to multicast[i] (an NVSwitch multicast mapping: one store lands on every GPU).  The
 * pointers address buffers with the same layout as x_dev (torch symmetric memory).  This fuses the
 * all-gather of the maps into the epilogue of the fit. */
#define MT_MAX_PEERS 8
#define MT_MAX_OWNED 32
typedef struct MtPeers {
    int32_t n_peers;
    int32_t n_owned; /* > 0: per-component destinations below are used (element-major x, n_comp <= 32) */
    float *peer[MT_MAX_PEERS];
    float *multicast;
    /* "component owner" exchange: component e of pixel p is additionally stored to owner_row[e][p]
     * (NULL: no extra store), a row of the buffer of the GPU that assembles the full map of e.  Every GPU
     * then receives only its own components from the others: a balanced all-to-all, (N-1)/N of the map
     * bytes per GPU instead of N-1 times at a single root.  With n_owned > 0 the final amplitudes are stored to
     * the owner rows ONLY (give a row for every component, the local GPU's own included); x_dev then just
     * carries the pixels handed to the non-negative solver. */
    float *owner_row[MT_MAX_OWNED];
} MtPeers;

/* Completion flags of the fused exchange: flags[r] addresses rank r's flag array (uint32 [n_ranks], zero
 * initialised, symmetric memory); rank q's signal of step k stores k to flags[r][q] on every r. */
typedef struct MtPeerSync {
    int32_t n_ranks;
    int32_t rank;
    uint32_t *flags[MT_MAX_PEERS];
} MtPeerSync;

int mt_abi_version(void);
const char *mt_last_error(void);

/* MT_OK iff the current CUDA device is a compute-capability-10.x part. */
int mt_device_check(void);

/* Energy axis of channels ch0 .. ch0 + n_ch - 1 on the device, ev = offset + slope * ch + quad * ch^2 with the rounding
 * of the reference's separate fp32 ops (map.py:271-281); channel numbers below 2^24 (exact in fp32). */
int mt_energy_axis(int32_t ch0, int32_t n_ch, float e_offset, float e_slope, float e_quad, float *ev_dev, void *stream);

/* K1 -- replaces the per-line torch op chains of model_elem_spec / elastic_peak /
 * compton_peak (map.py:72-124,152-257).  out_dev[col*ld_out + c] = sum of the column's
 * terms at channel c, fp32, accumulated in term order.  col_start_host has n_cols+1
 * entries indexing terms_host.  ev_dev: [n_ch] fp32 energy axis (map.py:271-281, built on
 * the host so that it is bit-identical to the reference's). */
int mt_model_terms(const float *ev_dev, int32_t n_ch, const MtTerm *terms_host,
                   const int32_t *col_start_host, int32_t n_cols, float *out_dev, int64_t ld_out,
                   void *stream);

/* Ordered column sum + optional escape-peak shift-add: replaces the agr_spec.add_ chain and
 * escape_peak of model_spec (map.py:282-308, 127-150).  out = sum_col cols[col] (in order);
 * if escape_bins > 0 and escape_factor > 0: out[c] += escape_factor * out_before[c + escape_bins]. */
int mt_model_sum(const float *cols_dev, int64_t ld_cols, int32_t n_cols, int32_t n_ch,
                 int32_t escape_bins, float escape_factor, float *out_dev, void *stream);

/* K2 -- replaces snip_bkg/snip_op (bkg.py:53-114) for a batch of spectra.
 * spec_dev: [n_px][ld_spec] of `dtype`, channels [ch0, ch0+n_ch) of every row are used.
 * lohi_dev: [n_pass][n_ch] uint32, lo | (hi << 16): the gather indices of each clipping pass,
 *           computed on the host with the reference's fp32 op sequence (bkg.py:57-62, 83-112).
 * out_mode: MT_SNIP_BACKGROUND  out_dev = fp32 [n_px][ld_out] background in counts
 *           MT_SNIP_RESIDUAL    out_dev = fp32 [n_px][ld_out] spectrum - background
 *           MT_SNIP_RESIDUAL_HILO out_dev = fp32 [n_px][ld_out], ld_out >= 2*n_ch: residual split
 *                               into a tf32-exact high part ([0,n_ch)) and remainder ([n_ch,2n_ch)) */
#define MT_SNIP_BACKGROUND 0
#define MT_SNIP_RESIDUAL 1
#define MT_SNIP_RESIDUAL_HILO 2
#define MT_SNIP_RESIDUAL_HILO_F16 3 /* mt_snip_residual_f16 only */
int mt_snip(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0,
            int32_t n_ch, const uint32_t *lohi_dev, int32_t n_pass, int32_t boxcar, float *out_dev,
            int64_t ld_out, int32_t out_mode, void *stream);

/* lohi_dev [n_pass][n_ch] of mt_snip built on the device from six fp32 calibration values with the reference's own
 * single-precision operation sequence (bkg.py:57-62, 83-91, 104-112): bit-identical to the host tables.  n_pass is the
 * pass count of bkg.py:104-112, which the host obtains from the largest width alone. */
int mt_snip_tables_f32(int32_t n_ch, int32_t ch0, float e_offset, float e_slope, float e_quad, float snip_width,
                       float fwhm_offset, float fwhm_fanoprime, int32_t n_pass, uint32_t *lohi_dev, void *stream);

/* The same kernel driven by a pre-scaled gather table: scaled_dev = (n_pass + 1) rows of ld_scaled words,
 * word = (lo * 16) | (hi * 16) << 16 (shared-memory byte offsets of the 4-pixel records), entries past n_ch and
 * the extra last row are 0 -- the clipping passes then need no index arithmetic and no bounds tests.
 * mt_snip_scaled_pitch(n_ch) gives the row pitch (n_ch rounded up to the kernel's thread grid; 0 when n_ch > 4096,
 * where byte offsets no longer fit 16 bits and mt_snip must be used); mt_snip_scale_tables converts lohi_dev
 * (once per calibration, one tiny launch). */
int mt_snip_scaled_pitch(int32_t n_ch);
int mt_snip_scale_tables(const uint32_t *lohi_dev, int32_t n_pass, int32_t n_ch, uint32_t *scaled_dev,
                         int32_t ld_scaled, void *stream);
int mt_snip_scaled(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0, int32_t n_ch,
                   const uint32_t *scaled_dev, int32_t ld_scaled, int32_t n_pass, int32_t boxcar, float *out_dev,
                   int64_t ld_out, int32_t out_mode, void *stream);

/* K2b -- the same computation (bit-identical results) for ALIGNED volumes, organised for throughput: contiguous channel
 * runs per thread (late passes take their neighbours from registers), persistent CTAs that fetch the next four rows by
 * cp.async.bulk while the current ones are written back, 128-bit global accesses (csrc/mt_snip_runs.cu).
 * mt_snip_runs_pitch(n_ch): row pitch in words of its table, 0 when n_ch is not served (n_ch % 8 != 0 or n_ch > 4096).
 * mt_snip_runs_tables: n_pass rows of ld_tabs words built from lohi_dev (gather words as swizzled shared-memory byte
 * offsets, then one descriptor per run of channels); the kernel stages one row per pass in shared memory.
 * mt_snip_runs_eligible: 1 when the kernel serves this problem: boxcar == 5, spec_dev / out_dev, the row pitches and the
 * crop start 16-byte aligned.  mt_snip_runs fails with MT_ERR_UNSUPPORTED otherwise (use mt_snip / mt_snip_scaled).
 * out_dev / ld_out / out_mode as in mt_snip, plus MT_SNIP_RESIDUAL_HILO_F16 (out_dev fp16, ld_out in fp16 elements,
 * overflow_dev as in mt_snip_residual_f16; may be NULL for the other modes). */
int mt_snip_runs_pitch(int32_t n_ch);
int mt_snip_runs_tables(const uint32_t *lohi_dev, int32_t n_pass, int32_t n_ch, uint32_t *tabs_dev, int32_t ld_tabs,
                        void *stream);
int mt_snip_runs_eligible(const void *spec_dev, int32_t dtype, int64_t ld_spec, int32_t ch0, int32_t n_ch, int32_t boxcar,
                          const void *out_dev, int64_t ld_out, int32_t out_mode);
int mt_snip_runs(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0, int32_t n_ch,
                 const uint32_t *tabs_dev, int32_t ld_tabs, int32_t n_pass, int32_t boxcar, void *out_dev, int64_t ld_out,
                 int32_t out_mode, int32_t *overflow_dev, void *stream);

/* K3 -- the per-pixel amplitude solve: replaces the Adam loop of fit_spec (opt.py:249-317)
 * for fixed global parameters, where the model is linear in the amplitudes.
 *   x[e][p] = colscale[e] * sum_d plane_ratio^d * ( sum_k spec[p][k] * wpack[d*e_pad + e][k] )
 * i.e. the spectra contracted on the tensor cores against the pseudo-inverse of the element
 * basis, which is stored as n_planes operand-precision "digit" planes stacked along N.  With
 * small-integer digits (10 bits per plane) and integer counts every partial sum is an integer
 * below 2^24, so the fp32 tensor-core accumulation is exact.
 * spec_dev : [n_px][ld_spec] of `dtype` (row pitch must be a multiple of 16 bytes, base 16-B aligned)
 * wpack_dev: [n_planes*e_pad][ld_w] in the SAME dtype as spec (fp32 words must hold tf32-exact
 *            values); columns outside the fitted channel range hold zeros.  e_pad multiple of 8,
 *            n_planes*e_pad a multiple of 16 and <= 256.
 * k_begin,k_end: channel window (elements) the kernel streams; rounded outwards to 128-byte blocks.
 * x_dev    : fp32 amplitudes in `x_layout` (MT_X_*), e < n_comp, p < n_px.
 * neg_list_dev / neg_count_dev (optional, both or neither): the epilogue appends px_offset + p of
 *            every pixel whose solution has a negative component to neg_list_dev (capacity >= n_px)
 *            and bumps *neg_count_dev (caller zeroes it); consumed by mt_nnls_list.
 * hinv32_dev (optional; needs e_pad <= 32, element-major x and the list): G^-1 in fp32 [n_comp][n_comp].
 *            The epilogue then applies the non-negative fix-up itself whenever a pixel's zero set has
 *            at most two components (closed form, registers only) and lists only the remaining pixels. */
int mt_fit_contract(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec,
                    int32_t n_ch_total, int32_t k_begin, int32_t k_end, const void *wpack_dev,
                    int64_t ld_w, int32_t e_pad, int32_t n_comp, int32_t n_planes,
                    const float *colscale_dev, float plane_ratio, float *x_dev, int64_t ld_x, int32_t x_layout,
                    int64_t px_offset, int32_t *neg_list_dev, int32_t *neg_count_dev,
                    const float *hinv32_dev, const MtPeers *peers_host /* optional */, void *stream);

/* Same contract on CUDA cores (no tensor cores, fp32 weights [n_comp][ld_w]); exists for
 * shapes/alignments TMA cannot address and as an on-device cross-check of mt_fit_contract. */
int mt_fit_contract_simt(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec,
                         int32_t k_begin, int32_t k_end, const float *w_dev, int64_t ld_w,
                         int32_t n_comp, float *x_dev, int64_t ld_x, int32_t x_layout, void *stream);

/* Batched per-pixel non-negative fix-up of the least-squares amplitudes: minimise
 * (x-xu)^T G (x-xu) s.t. x >= 0 with G = B^T B shared by all pixels (the converged answer of
 * the reference's 10**a parametrisation, opt.py:154).  g_dev = G and hinv_dev = G^-1, both
 * fp64 [n_comp][n_comp] row-major, n_comp <= 104.  Pixels whose unconstrained solution is
 * already non-negative are untouched.  n_fixed_dev (optional, may be NULL): int32 counter,
 * incremented once per pixel that needed the fix-up. */
int mt_nnls_fixup(float *x_dev, int64_t ld_x, int32_t x_layout, int64_t n_px, int32_t n_comp, const double *g_dev,
                  const double *hinv_dev, int32_t *n_fixed_dev, const MtPeers *peers_host /* optional */,
                  void *stream);

/* Amplitude-sparsity penalty of the MSE objective -- replaces the `l1_lambda` branch of _calculate_loss
 * (opt.py:372-396) for the map fit.  With A >= 0 the penalty c*sum(A) is linear, so it only shifts the
 * unconstrained amplitudes:  x[e,p] -= coef[e] * S_p,  S_p = sum_c mask[c] * (sum_{r<repeat} spec[p, col0 + r*n_ch + c])^2
 * (the row's squared distance from the background, i.e. loss(bkg, y) of the reference), and
 * coef[e] = l1_lambda / (2 * n_elements) * (G^-1 1)[e] is prepared by the host.  `repeat` = 2 reads the
 * exact hi + remainder operand pair the SNIP kernel writes.  chan_mask_dev may be NULL (all channels).
 * Follow with mt_nnls_fixup. */
int mt_penalty_shift(const void *spec_dev, int32_t spec_dtype, int64_t n_px, int64_t ld_spec, int32_t col0,
                     int32_t n_ch, int32_t repeat, const float *chan_mask_dev, const float *coef_dev, float *x_dev,
                     int64_t ld_x, int32_t x_layout, int32_t n_comp, void *stream);

/* Exact operand split for fp32 volumes whose values need more than the 11 significant bits of a tf32 operand
 * (integer counts >= 2048): out[p][0..n_ch) = spec[p][col0 + c] rounded to nearest to 11 significant bits,
 * out[p][n_ch..2 n_ch) = the remainder (exact).  mt_fit_contract then runs over the 2 n_ch columns with the
 * operand pack repeated twice -- the layout MT_SNIP_RESIDUAL_HILO produces for the SNIP path. */
int mt_split_hilo(const float *spec_dev, int64_t n_px, int64_t ld_spec, int32_t col0, int32_t n_ch, float *out_dev,
                  int64_t ld_out, void *stream);

/* K2 writing the residual as TWO FP16 PLANES for the fp16 tensor-core path of mt_fit_contract: out_half_dev =
 * fp16 [n_px][ld_out] (ld_out in fp16 elements, >= 2*n_ch): [0,n_ch) = spectrum - background rounded to fp16,
 * [n_ch,2n_ch) = the remainder (22 significant bits together, like MT_SNIP_RESIDUAL_HILO, at half the bytes: the
 * residual crosses HBM as 4 bytes per channel instead of 8).  *overflow_dev |= 1 (int32, zeroed by the caller) when a
 * residual exceeds the fp16 range (|r| > 65504); the caller then repeats that slab with MT_SNIP_RESIDUAL_HILO.
 * tab_dev / ld_tab: the gather table of mt_snip (scaled = 0, ld_tab ignored) or of mt_snip_scaled (scaled = 1). */
int mt_snip_residual_f16(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0, int32_t n_ch,
                         const uint32_t *tab_dev, int32_t ld_tab, int32_t scaled, int32_t n_pass, int32_t boxcar,
                         void *out_half_dev, int64_t ld_out, int32_t *overflow_dev, void *stream);

/* Operand-exactness test for fp32 volumes on the tensor-core path: *flag_dev |= 1 (int32, zeroed by the caller)
 * when any value of spec[p][col0 .. col0+n_ch) needs more than the 11 significant bits of a tf32 operand (or is
 * not finite).  The reference accepts arbitrary float spectra (opt.py:344-375); volumes that fail the test are
 * routed through mt_split_hilo. */
int mt_operand_inexact(const float *spec_dev, int64_t n_px, int64_t ld_spec, int32_t col0, int32_t n_ch,
                       int32_t *flag_dev, void *stream);

/* Number of pixels, since the last reset, whose non-negative pivot system was numerically singular (near-collinear
 * components); such pixels receive the clamped unconstrained solution.  Synchronises the device. */
int mt_nnls_singular_count(uint64_t *count_host, int32_t reset);

/* Same solve, driven by the pixel list mt_fit_contract compacted (no scan of x, every lane busy):
 * list_dev[i] for i < *count_dev are pixel numbers p (x_dev[e*ld_x + p]).  short_list != 0 selects
 * the warp-per-pixel variant (n_comp <= 32, element-major), which minimises the latency of a solve
 * and is the right choice when the K3 epilogue has already fixed most pixels. */
int mt_nnls_list(float *x_dev, int64_t ld_x, int32_t x_layout, const int32_t *list_dev, const int32_t *count_dev,
                 int32_t short_list,
                 int32_t n_comp, const double *g_dev, const double *hinv_dev,
                 const MtPeers *peers_host /* optional */, void *stream);

/* Channel-first slabs (raw XRF-Maps MAPS/Spectra/mca_arr is [C][H][W]; create_dataset moves the energy axis last on
 * the host, io.py:154-156): dst[p * ld_dst + c] = src[c * ld_src + p] for c < n_ch, p < n_px, same dtype. */
int mt_transpose_cf(const void *src_dev, int32_t dtype, int32_t n_ch, int64_t n_px, int64_t ld_src, void *dst_dev,
                    int64_t ld_dst, void *stream);

/* Integrated spectrum (MAPS/int_spec, io.py:159): acc_dev[c] += sum_p spec[p * ld_spec + c], fp64 accumulators
 * (exact and order-independent for integer counts), accumulated over successive calls (slabs); the caller zeroes
 * acc_dev first. */
int mt_int_spec(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t n_ch, double *acc_dev,
                void *stream);

/* Per-pixel loss="l1" maps (opt.py:232-233, 371-375) by ADMM on z = model - target; see csrc/mt_lad.cu.  For every
 * pixel p and channel c < n_ch (row pitches ld_*, element-major x_dev [n_comp][ld_x], basis_dev [n_comp][ld_b] fp32):
 *   m = sum_e x[e][p] basis[e][c],  r = m - t[p][c],  obj_dev[p] = sum_c |r|   (optional output: the L1 objective of x)
 * and, with mode = 1:  v = r + u;  z = soft(v, kappa_dev[p]);  u <- v - z;  target = t + z - u;  dq = target -
 * qstate[p][c];  qstate <- target;  q[p][c], q[p][n_ch + c] <- dq split into a tf32-exact high part and remainder -- the
 * operand of the next x-update: xu += mt_fit_contract(dq) (2 n_ch columns, pack repeated twice, no fix-up), then the
 * non-negative fix-up of a copy of xu.  qstate starts as t with xu the least-squares solution of t.  mode = 0 only
 * evaluates the objective. */
int mt_lad_update(const float *t_dev, int64_t ld_t, const float *x_dev, int64_t ld_x, const float *basis_dev, int64_t ld_b,
                  int32_t n_comp, int32_t n_ch, int64_t n_px, float *u_dev, int64_t ld_u, const float *kappa_dev,
                  float *qstate_dev, int64_t ld_qs, float *q_dev, int64_t ld_q, float *obj_dev, int32_t mode, void *stream);

/* The collective that is left of the map gather once the stores are fused into the fit:
 * mt_peer_signal increments *epoch_dev and publishes it to every rank's flag array after everything queued on
 * `stream` so far (the fit kernels' peer stores) has completed; mt_peer_wait blocks the stream until every rank
 * has published an epoch >= *epoch_dev - lag into this rank's array, i.e. until this rank holds every peer's share
 * of the maps of that step.  lag = 1 after the signal of step k waits for everybody's step k-1: with double-buffered
 * maps the ranks then run with one step of slack and the barrier of step k overlaps step k+1.  The wait
 * is bounded (~10 s): on expiry *timed_out_dev (optional int32) is set to 1 + the silent rank. */
int mt_peer_signal(const MtPeerSync *sync_host, uint32_t *epoch_dev, void *stream);
int mt_peer_wait(const MtPeerSync *sync_host, const uint32_t *epoch_dev, int32_t lag, int32_t *timed_out_dev,
                 void *stream);

/* ---- K4: batched per-pixel nonlinear refinement (BASELINE config 4) ----------------------------
 * Replaces fit_spec(..., fitting_params=[...], tune_params=True) (opt.py:165-327) applied per pixel:
 * every pixel refines up to four calibration parameters together with all amplitudes by a damped
 * Gauss-Newton iteration with analytic Jacobians on sum_c (model_c + bkg_c - y_c)^2.
 * Model terms = model_spec's path (map.py:260-309: use_step=True, use_tail=False), per line
 *   Gaussian  x[comp] * factor * slope / (sg sqrt(2 pi)) * exp(-(ev_c - energy)^2 / (2 sg^2))          (map.py:57-58, 208)
 *   step      x[comp] * factor * step * slope / (2 energy) * erfc((ev_c - energy) / (sqrt(2) sb))     (map.py:61-62, 209-214)
 * with sb = sqrt((FWHM_OFFSET / 2.3548)^2 + energy * 2.96 * FWHM_FANOPRIME), sg = sigma_scale * sb, and the energy of
 * the two scatter lines following COHERENT_SCT_ENERGY / COMPTON_ANGLE (map.py:72-124).  The free parameters are the
 * eight calibration parameters with a non-zero gradient on that path (the tail fractions of default_fitting_params,
 * default.py:51-64, do not enter model_spec: use_tail=False at map.py:290,302). */
#define MT_THETA_ENERGY_OFFSET 0
#define MT_THETA_ENERGY_SLOPE 1
#define MT_THETA_FWHM_OFFSET 2
#define MT_THETA_FWHM_FANOPRIME 3
#define MT_THETA_ENERGY_QUADRATIC 4
#define MT_THETA_COHERENT_SCT_ENERGY 5
#define MT_THETA_COMPTON_ANGLE 6
#define MT_THETA_COMPTON_FWHM_CORR 7
#define MT_N_THETA_IDS 8
#define MT_MAX_THETA 8        /* free parameters of mt_model_jacobian */
#define MT_MAX_THETA_REFINE 4 /* free parameters per pixel in mt_refine */

#define MT_LINE_ELEMENT 0
#define MT_LINE_ELASTIC 1 /* energy = COHERENT_SCT_ENERGY */
#define MT_LINE_COMPTON 2 /* energy = Compton energy of COHERENT_SCT_ENERGY at COMPTON_ANGLE; Gaussian width x COMPTON_FWHM_CORR */

typedef struct MtRefineLine {
    float energy;      /* keV at the global parameters (scatter lines follow the free parameters) */
    float factor;      /* branching ratio / normalisation of the line (amplitude excluded): ratio / (1 + f_step) */
    int32_t comp;      /* component (row of x) the line belongs to */
    float sigma_scale; /* 1, or COMPTON_FWHM_CORR for the Compton Gaussian (map.py:105) */
    float e296;        /* energy * 2.96 as the reference rounds it (map.py:167) */
    float step;        /* f_step of the line (map.py:169-172; COMPTON_F_STEP for the Compton line); 0 = no step term */
    int32_t kind;      /* MT_LINE_* */
    int32_t reserved;
} MtRefineLine;

typedef struct MtRefineConfig {
    int32_t n_ch, ch0;  /* fitted channels [ch0, ch0 + n_ch) of each spectrum row */
    int32_t n_comp, n_lines, n_theta, max_iter;
    int32_t theta_id[MT_MAX_THETA]; /* MT_THETA_* of each free parameter */
    float theta0[MT_MAX_THETA], theta_lo[MT_MAX_THETA], theta_hi[MT_MAX_THETA];
    float energy_offset, energy_slope, energy_quad, fwhm_offset, fwhm_fanoprime; /* global values */
    float coherent_sct_energy, compton_angle, compton_fwhm_corr;
    float tol;          /* stop when steps fall below tol (relative) */
} MtRefineConfig;

/* Model of one spectrum and its analytic derivatives w.r.t. the free parameters at theta0, amplitudes x_dev [n_comp]
 * fixed -- what torch.autograd gives the reference for d model_spec / d param (opt.py:249-317), and the Jacobian of
 * fit_spec(tune_params=True) here.  model_dev [n_ch], dtheta_dev [n_theta][ld] (ld >= n_ch), up to MT_MAX_THETA
 * parameters.  lines_host / chan_lines_host as for mt_refine. */
int mt_model_jacobian(const MtRefineConfig *cfg_host, const MtRefineLine *lines_host, const uint32_t *chan_lines_host,
                      const float *x_dev, float *model_dev, float *dtheta_dev, int64_t ld, void *stream);

/* lines_host sorted by energy; chan_lines_host[c] = lo | hi << 16 is the range of line indices that can
 * contribute at channel c; comp_start_host/comp_lines_host list each component's lines;
 * line_window_host[l] = first | (last + 1) << 16 local channel of line l's window; comp_order_host
 * [n_comp] is a permutation of the components: warp w of the CTA's 4 takes entries w, w + 4, ... in the
 * component-major pass, so the host orders them to balance the summed window lengths; g0inv_dev = inverse
 * Gram matrix of the unit-amplitude basis at the global parameters, fp32 [n_comp][n_comp].
 * x_dev [n_comp][ld_x] holds the starting amplitudes (from mt_fit_contract) and receives the result;
 * theta_dev [n_theta][ld_theta], loss_dev [n_px], iters_dev [n_px] are optional outputs.
 * chan_weight_dev (optional, fp32 [n_ch], 0 or 1): the channels that enter the objective -- `indices` of fit_spec
 * (opt.py:344-347); g0inv_dev must then be the inverse Gram matrix of the masked basis.
 * escape_bins > 0 with escape_factor > 0: every component carries its detector escape peak, b'(c) = b(c) + escape_factor *
 * b(c + escape_bins) for c + escape_bins < n_ch (escape_peak, map.py:127-150; the bin shift is the one of the global
 * calibration); g0inv_dev then belongs to that augmented basis. */
int mt_refine(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, const float *bkg_dev,
              int64_t ld_bkg, const float *chan_weight_dev, int32_t escape_bins, float escape_factor,
              const MtRefineConfig *cfg_host, const MtRefineLine *lines_host,
              const uint32_t *chan_lines_host, const int32_t *comp_start_host, const int32_t *comp_lines_host,
              const uint32_t *line_window_host, const int32_t *comp_order_host, const float *g0inv_dev,
              float *x_dev, int64_t ld_x,
              float *theta_dev, int64_t ld_theta, float *loss_dev, int32_t *iters_dev, void *stream);

/* Amplitude solve of ONE spectrum at fixed calibration -- the evaluation inside fit_spec(tune_params=True)
 * (objective of opt.py:329-398; the reference reaches the same amplitudes by Adam on log10 A).  For each of n_prob
 * problems (the trial points of one Jacobian), over the n_k fitted channels c_k = chan_idx_dev[k] (NULL: c_k = k):
 *     x = argmin_{x >= 0} sum_k w_k (x . B[:, c_k] - t_k)^2 + 2 shift sum(x),   t = y - bkg (fp32 difference),
 * with c = l1_lambda * (robust ? sum|t| : sum t^2) / max(n_elements, 1) (opt.py:372-381) and shift = robust ? c : c/2.
 * basis_dev fp32 [n_prob][n_comp][ld_b] (problem p at p * prob_stride_b; 0 shares one basis), y_dev [>= max c_k + 1],
 * bkg_dev optional (problem p at p * ld_bkg; ld_bkg = 0 shares one), weights_dev optional fp64 [n_prob][ld_w] indexed by k.
 * Outputs, all fp64: x_dev [n_prob][n_comp]; res_dev [n_prob][n_k] = model - t; vec_dev [n_prob][n_k + n_comp] =
 * sqrt(w) * res followed by the penalty entries sqrt((robust ? 2 : 1) * c * x); stats_dev [n_prob][8] = { |vec|^2,
 * objective (sum res^2 or sum |res|, + c sum x), c, flag (0 ok, 1 singular pivot block: clamped unconstrained solution,
 * 2 Gram matrix of the live components singular: x = 0), pivot steps, sum res^2, sum |res|, sum x }.
 * free_set_dev (optional, [n_prob][n_comp] bytes, in / out): the partition the pivoting starts from (1 free, 0 held at zero,
 * any other value: no hint) and, on return, the final one (2 = component without signal) -- the warm start between
 * neighbouring trial points.  n_comp <= MT_SPEC_MAX_COMP, n_k <= MT_SPEC_MAX_CH.  One launch, no host synchronisation. */
#define MT_SPEC_MAX_COMP 96
#define MT_SPEC_MAX_CH 8192
int mt_spec_solve(const float *basis_dev, int64_t ld_b, int64_t prob_stride_b, int32_t n_comp, const float *y_dev,
                  const float *bkg_dev, int64_t ld_bkg, const int32_t *chan_idx_dev, int32_t n_k,
                  const double *weights_dev, int64_t ld_w, double l1_lambda, int32_t robust, int32_t n_elements,
                  double *x_dev, double *res_dev, double *vec_dev, double *stats_dev, uint8_t *free_set_dev, int32_t n_prob,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MAPSTORCH_B200_H */
