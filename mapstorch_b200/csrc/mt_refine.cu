// K4 -- batched per-pixel nonlinear refinement (BASELINE config 4).
//
// The reference can free calibration parameters per spectrum with
// fit_spec(..., fitting_params=[...], tune_params=True) (opt.py:201,119,249-317): Adam over autograd
// on the objective sum_c (model_c(theta, a) + bkg_c - y_c)^2.  Here every pixel refines up to four
// peak-shape parameters theta (energy offset / slope, FWHM offset / Fano term) together with all
// amplitudes by a damped Gauss-Newton iteration with ANALYTIC Jacobians:
//
//   one CTA (128 threads) per pixel, per iteration
//   B1  channel-major : every thread evaluates its channels: model m_c, the theta-derivatives
//                       d_k(c) = dm_c/dtheta_k and the residual r_c (kept in shared memory), with
//                       thread-local partial sums of the loss, D^T r and D^T D
//   B2  component-major: every warp owns whole components and reduces B^T r and D^T B over the
//                       line windows (deterministic: no atomics, fixed reduction trees)
//   C   solve         : the amplitude block of the Gauss-Newton matrix is the pixel-independent
//                       Gram matrix G0 (its inverse is an input), the theta block and the coupling
//                       are exact, so the step is a Schur complement on an n_theta x n_theta system
//
// The gradient is exact, only the amplitude curvature is frozen at the global parameters, so the
// fixed point is the true minimiser.  Model terms are those of model_spec's path (map.py:260-309): the
// Gaussian of every line, elastic and Compton peaks, and -- GENERAL instantiations -- the step terms
// (map.py:61-62, 209-214, 111-114) and the scatter lines whose energy / width follow COHERENT_SCT_ENERGY,
// COMPTON_ANGLE and COMPTON_FWHM_CORR.  Tail terms never enter model_spec (use_tail=False, map.py:290,302).
// exp dominates (MUFU bound).  mt_model_jacobian exposes the same evaluation for one spectrum (model and
// d model / d theta): the analytic Jacobian of fit_spec(tune_params=True).
#include <cuda_fp16.h>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "mt_common.cuh"
#include "mt_f32x2.cuh"

namespace {

using mt::add2;
using mt::bcast2;
using mt::f32x2;
using mt::fma2;
using mt::hsum2;
using mt::mul2;
using mt::pack2;
using mt::unpack2;

constexpr int kRefThreads = 128;  // 256 threads per pixel measured the same (70.1 ms) at equal warps per SM
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kMaxLines = 512;
constexpr int kMaxComp = 64;
constexpr int kMaxTheta = MT_MAX_THETA_REFINE;
constexpr int kRefMinBlocks = 7;  // round 1 (c4): 4 -> 77.0 ms, 5 -> 72.3, 6 -> 69.6, 7/8 -> 70.0; round 2 kernel: 6 -> 30.6 ms, 7 -> 29.6 (72 registers, no spills)

struct RefParams {
    MtRefineConfig cfg;
    const MtRefineLine *lines;      // [n_lines], sorted by energy
    const uint32_t *chan_lines;     // [n_ch]: lo | hi << 16, line index range covering the channel
    const uint32_t *pair_info;      // [ceil(n_ch / 2)]: channel pair | first line << 12 | line count << 22, sorted by count
    const uint32_t *b2_items;       // component-major pass, flattened: line | channel pair << 10 (line 0x3ff = padding); NULL:
    const int32_t *b2_start;        // [n_comp + 1] item offsets (multiples of 32)      per-line window loops instead
    const int32_t *comp_start;      // [n_comp + 1] into comp_lines
    const int32_t *comp_lines;      // line indices grouped by component
    const uint32_t *line_window;    // [n_lines]: first | last+1 << 16 channel (local) of the line's window
    const int32_t *comp_order;      // [n_comp]: permutation; warp w takes entries w, w + 4, ... (balanced by the host)
    const float *g0inv;             // [n_comp][n_comp]
    const float *bkg;
    long long ld_bkg;
    const float *chan_w;            // optional [n_ch]: weight (0 / 1) of each channel in the objective (fit_spec's `indices`)
    int escape_bins;                // > 0: every column carries its escape peak, b'(c) = b(c) + escape_factor * b(c + bins)
    float escape_factor;            //      (map.py:127-150); GENERAL instantiations only
    const void *spec;
    long long ld_spec;
    float *x;
    long long ld_x;
    float *theta_out;
    long long ld_theta;
    float *loss_out;
    int *iters_out;
    long long n_px;
};

template <typename T>
__device__ __forceinline__ float ld_spec(const T *p);
template <>
__device__ __forceinline__ float ld_spec<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_spec<__half>(const __half *p) { return __half2float(__ldg(p)); }

// 2^x on the SFU (ex2.approx.ftz: 2 ulp, results below 2^-126 flush to 0 -- they are line tails 38 sigma^2 out)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// ---- calibration state and per-line constants ---------------------------------------------------------------
struct Calib {
    float off, slope, quad, fo, fa, e0, angle, corr;
};

__device__ __forceinline__ Calib calib_of(const MtRefineConfig &cfg, const float *theta, int n_theta) {
    Calib c{cfg.energy_offset, cfg.energy_slope, cfg.energy_quad, cfg.fwhm_offset, cfg.fwhm_fanoprime,
            cfg.coherent_sct_energy, cfg.compton_angle, cfg.compton_fwhm_corr};
    for (int k = 0; k < n_theta; ++k) {
        const float v = theta[k];
        switch (cfg.theta_id[k]) {
            case MT_THETA_ENERGY_OFFSET: c.off = v; break;
            case MT_THETA_ENERGY_SLOPE: c.slope = v; break;
            case MT_THETA_FWHM_OFFSET: c.fo = v; break;
            case MT_THETA_FWHM_FANOPRIME: c.fa = v; break;
            case MT_THETA_ENERGY_QUADRATIC: c.quad = v; break;
            case MT_THETA_COHERENT_SCT_ENERGY: c.e0 = v; break;
            case MT_THETA_COMPTON_ANGLE: c.angle = v; break;
            case MT_THETA_COMPTON_FWHM_CORR: c.corr = v; break;
        }
    }
    return c;
}

// Phase A for one line.  Shared-memory records (one 128-bit load each in the inner loops):
//   lc0 = {energy, a = -log2(e) / (2 sg^2), amplitude * unit, window (first | last+1 << 16)}
//   lc1 = {1 / sg^2, (d sg / d FWHM_OFFSET) / sg, (d sg / d FWHM_FANOPRIME) / sg, unit}
// GENERAL only:
//   lc2 = {amplitude * K, -log2(e) / (2 sb^2), 1 / (sqrt(2) sb), amplitude * K * 2 / sqrt(pi) / (sqrt(2) sb)}   (step term)
//   lc3 = {K, 1 / energy, 0, 0}
//   lck[k * L + l] = {(d sg / d theta_k) / sg, d energy / d theta_k, (d sb / d theta_k) / sb, 0}  (0 for the axis parameters)
// with unit = factor * slope / (sg sqrt(2 pi)), K = factor * step * slope / (2 energy), sb the line width, sg = scale * sb.
template <int NT, bool GENERAL>
__device__ __forceinline__ void line_constants(const MtRefineConfig &cfg, const Calib &c, const MtRefineLine &ln, float amp,
                                               uint32_t window, int l, int L, float4 *lc0, float4 *lc1, float4 *lc2,
                                               float4 *lc3, float4 *lck) {
    const float s0 = c.fo / 2.3548f;
    float energy = ln.energy, e296 = ln.e296, scale = ln.sigma_scale, de_e0 = 0.f, de_ang = 0.f;
    if (GENERAL && ln.kind != MT_LINE_ELEMENT) {
        if (ln.kind == MT_LINE_ELASTIC) {
            energy = c.e0;
            de_e0 = 1.f;
        } else {
            const float rad = c.angle * 0.017453292519943295f;
            const float aa = (c.e0 / 511.f) * (1.f - cosf(rad));
            const float inv = 1.f / (1.f + aa);
            energy = c.e0 * inv;
            de_e0 = inv * inv;
            de_ang = -c.e0 * inv * inv * (c.e0 / 511.f) * sinf(rad) * 0.017453292519943295f;
            scale = c.corr;
        }
        e296 = energy * 2.96f;
    }
    const float sb = sqrtf(fmaxf(s0 * s0 + e296 * c.fa, 1e-12f));
    const float sg = sb * scale;
    const float unit = ln.factor * c.slope / (sg * 2.5066282746310002f);
    const float is = 1.f / sg, is2 = is * is, isb2 = 1.f / (sb * sb);
    const float c_fo = s0 * isb2 / 2.3548f;   // (d sb / d FWHM_OFFSET) / sb  (= the same for sg)
    const float c_fa = e296 * isb2 * 0.5f;    // (d sb / d FWHM_FANOPRIME) / sb
    lc0[l] = make_float4(energy, -0.72134752044448170f * is2, unit * amp, __uint_as_float(window));
    lc1[l] = make_float4(is2, c_fo, c_fa, unit);
    if (GENERAL) {
        const float K = ln.step > 0.f ? ln.factor * ln.step * c.slope / (2.f * energy) : 0.f;
        const float zs = 0.70710678118654752f / sb;
        lc2[l] = make_float4(amp * K, -0.72134752044448170f * isb2, zs, amp * K * 1.1283791670955126f * zs);
        lc3[l] = make_float4(K, 1.f / energy, 0.f, 0.f);
        const float c_e = 1.48f * c.fa * isb2;  // (d sb / d energy) / sb
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            float csb = 0.f, csg = 0.f, ce = 0.f;
            switch (cfg.theta_id[k]) {
                case MT_THETA_FWHM_OFFSET: csb = csg = c_fo; break;
                case MT_THETA_FWHM_FANOPRIME: csb = csg = c_fa; break;
                case MT_THETA_COHERENT_SCT_ENERGY: ce = de_e0; csb = csg = c_e * de_e0; break;
                case MT_THETA_COMPTON_ANGLE: ce = de_ang; csb = csg = c_e * de_ang; break;
                case MT_THETA_COMPTON_FWHM_CORR: csg = ln.kind == MT_LINE_COMPTON ? 1.f / c.corr : 0.f; break;
                default: break;  // axis parameters: handled after the line loop
            }
            lck[k * L + l] = make_float4(csg, ce, csb, 0.f);
        }
    }
}

__device__ __forceinline__ void erfc2(f32x2 z, float &lo, float &hi) {
    float a, b;
    unpack2(z, a, b);
    lo = erfcf(a);
    hi = erfcf(b);
}

// Model and theta derivatives of ONE PAIR of adjacent channels over the lines [l_begin, l_end): packed fp32 arithmetic
// (FADD2 / FMUL2 / FFMA2: two channels per issue slot, the per-line constants enter as broadcast operands).
// chan_lines ranges are unions of line windows, so a few (line, channel) pairs outside the line's own window are
// evaluated too -- they are < e^-15 of the peak.
template <int NT, bool GENERAL>
__device__ __forceinline__ void eval_pair(const int (&id_k)[NT > 0 ? NT : 1], const Calib &c, f32x2 cg, int l_begin, int l_end,
                                          int L, const float4 *lc0, const float4 *lc1, const float4 *lc2, const float4 *lck,
                                          f32x2 &m_out, f32x2 (&d)[NT > 0 ? NT : 1]) {
    const f32x2 ev = fma2(fma2(bcast2(c.quad), cg, bcast2(c.slope)), cg, bcast2(c.off));
    const f32x2 minus_one2 = bcast2(-1.f);
    f32x2 m = 0ull, d_ev = 0ull;
    if (!GENERAL) {
        f32x2 d_fo = 0ull, d_fa = 0ull;
        for (int l = l_begin; l < l_end; ++l) {
            const float4 a0 = lc0[l];
            const float4 a1 = lc1[l];
            const f32x2 dl = add2(ev, bcast2(-a0.x)), dl2 = mul2(dl, dl);
            float e0, e1;
            unpack2(mul2(dl2, bcast2(a0.y)), e0, e1);
            const f32x2 v = mul2(pack2(ex2_approx(e0), ex2_approx(e1)), bcast2(a0.z));  // amplitude * unit * exp(-u^2/2)
            m = add2(m, v);
            d_ev = fma2(mul2(v, dl), bcast2(-a1.x), d_ev);                 // -v u / sigma
            const f32x2 w = mul2(v, fma2(dl2, bcast2(a1.x), minus_one2));  // v (u^2 - 1); 1 / sigma sits in a1.y, a1.z
            d_fo = fma2(w, bcast2(a1.y), d_fo);
            d_fa = fma2(w, bcast2(a1.z), d_fa);
        }
        // which of the three accumulated derivatives belongs to parameter k is the same for every pair of the launch:
        // two multiplications by 0 / 1 instead of chains of selects on 64-bit register pairs (ncu r02_refine_v4:
        // ~50 of the 106 instructions per pair were selects)
#pragma unroll
        for (int k = 0; k < NT; ++k)
            d[k] = fma2(d_fo, bcast2(id_k[k] == MT_THETA_FWHM_OFFSET ? 1.f : 0.f),
                        mul2(d_fa, bcast2(id_k[k] == MT_THETA_FWHM_FANOPRIME ? 1.f : 0.f)));
    } else {
#pragma unroll
        for (int k = 0; k < NT; ++k) d[k] = 0ull;
        for (int l = l_begin; l < l_end; ++l) {
            const float4 a0 = lc0[l];
            const float4 a1 = lc1[l];
            const f32x2 dl = add2(ev, bcast2(-a0.x)), dl2 = mul2(dl, dl);
            float e0, e1;
            unpack2(mul2(dl2, bcast2(a0.y)), e0, e1);
            const f32x2 v = mul2(pack2(ex2_approx(e0), ex2_approx(e1)), bcast2(a0.z));
            m = add2(m, v);
            const f32x2 t1 = mul2(mul2(v, dl), bcast2(a1.x));              // -d v / d ev = d v / d energy
            d_ev = add2(d_ev, mul2(t1, minus_one2));
            const f32x2 t2 = mul2(v, fma2(dl2, bcast2(a1.x), minus_one2));  // sg * d v / d sg
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float4 ck = lck[k * L + l];
                d[k] = fma2(t2, bcast2(ck.x), fma2(t1, bcast2(ck.y), d[k]));
            }
            const float4 a2 = lc2[l];
            if (a2.x != 0.f) {
                // step term: s = A K erfc(z), z = dl / (sqrt(2) sb);  -d s / d ev = q = A K 2/sqrt(pi) exp(-z^2) / (sqrt(2) sb)
                float r0, r1, g0, g1;
                erfc2(mul2(dl, bcast2(a2.z)), r0, r1);
                unpack2(mul2(dl2, bcast2(a2.y)), g0, g1);
                const f32x2 sv = mul2(pack2(r0, r1), bcast2(a2.x));
                const f32x2 q = mul2(pack2(ex2_approx(g0), ex2_approx(g1)), bcast2(a2.w));
                m = add2(m, sv);
                d_ev = add2(d_ev, mul2(q, minus_one2));
                const f32x2 qdl = mul2(q, dl);
                const f32x2 qe = fma2(sv, bcast2(-1.f / a0.x), q);  // d s / d energy = q - s / energy
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const float4 ck = lck[k * L + l];
                    d[k] = fma2(qdl, bcast2(ck.z), fma2(qe, bcast2(ck.y), d[k]));
                }
            }
        }
    }
    // the energy-axis parameters act through ev alone (+ the gain factor of ENERGY_SLOPE, map.py:57-62)
    // d[k] <- keep * d[k] + d_ev * (s0 + cg * (s1 + cg * s2)) + m * sm with launch-uniform 0 / 1 (and 1 / slope) factors
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        const int id = id_k[k];
        const bool axis = id == MT_THETA_ENERGY_OFFSET || id == MT_THETA_ENERGY_SLOPE || id == MT_THETA_ENERGY_QUADRATIC;
        const float s0 = id == MT_THETA_ENERGY_OFFSET ? 1.f : 0.f, s1 = id == MT_THETA_ENERGY_SLOPE ? 1.f : 0.f,
                    s2 = id == MT_THETA_ENERGY_QUADRATIC ? 1.f : 0.f, sm = id == MT_THETA_ENERGY_SLOPE ? 1.f / c.slope : 0.f;
        const f32x2 poly = fma2(fma2(bcast2(s2), cg, bcast2(s1)), cg, bcast2(s0));
        d[k] = fma2(d_ev, poly, fma2(m, bcast2(sm), mul2(d[k], bcast2(axis ? 0.f : 1.f))));
    }
    m_out = m;
}

template <typename T, int NT, bool GENERAL>
__global__ void __launch_bounds__(kRefThreads, GENERAL ? 4 : kRefMinBlocks)
refine_kernel(const RefParams p) {
    extern __shared__ float4 dyn4[];  // lc0[L], lc1[L], (GENERAL: lc2[L], lc3[L], lck[NT][L]), then r[n_ch], d[n_theta][n_ch]
    const MtRefineConfig &cfg = p.cfg;
    const int n_ch = cfg.n_ch, E = cfg.n_comp, L = cfg.n_lines;  // NT (free parameters) is a template argument
    float4 *lc0 = dyn4;
    float4 *lc1 = dyn4 + L;
    float4 *lc2 = dyn4 + 2 * L, *lc3 = dyn4 + 3 * L, *lck = dyn4 + 4 * L;
    const int n_rec = GENERAL ? (4 + NT) * L : 2 * L;
    const int ld = (n_ch + 1) & ~1;  // even pitch: channel pairs are read and written with 64-bit accesses
    float *r_s = reinterpret_cast<float *>(dyn4 + n_rec);
    float *d_s = r_s + ld;  // d_s[k] = r_s + (k + 1) * ld
    __shared__ float xs[kMaxComp], x_prev[kMaxComp], dx[kMaxComp], gx[kMaxComp], w_s[kMaxComp];
    __shared__ float hthx[kMaxTheta][kMaxComp], t1[kMaxComp][kMaxTheta];
    __shared__ float part[kRefWarps][1 + kMaxTheta + kMaxTheta * kMaxTheta];
    __shared__ float theta[kMaxTheta], theta_prev[kMaxTheta], dtheta[kMaxTheta];
    __shared__ float ctl[4];  // loss, loss_prev, alpha, flag

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long px = blockIdx.x;
    const T *row = static_cast<const T *>(p.spec) + px * p.ld_spec + cfg.ch0;
    const float *brow = p.bkg ? p.bkg + px * p.ld_bkg : nullptr;

    for (int e = tid; e < E; e += kRefThreads) {
        xs[e] = fmaxf(p.x[(long long)e * p.ld_x + px], 0.f);
        x_prev[e] = xs[e];
        dx[e] = 0.f;
    }
    if (tid < kMaxTheta) {
        theta[tid] = tid < NT ? cfg.theta0[tid] : 0.f;
        theta_prev[tid] = theta[tid];
        dtheta[tid] = 0.f;
    }
    if (tid == 0) {
        ctl[1] = 3.0e38f;  // loss_prev
        ctl[2] = 1.f;      // alpha
    }
    __syncthreads();

    // largest active set the scratch (residual + derivative rows) can hold: cap^2 + cap (NT + 1) floats
    int cap = E;
    while (cap > 0 && cap * cap + cap * (NT + 1) > (1 + NT) * ld) --cap;

    int iter = 0;
    for (;; ++iter) {
        // ---- current calibration ------------------------------------------------------------
        const Calib cal = calib_of(cfg, theta, NT);
        const float off = cal.off, slope = cal.slope;
        // ---- A: per-line shape constants ---------------------------------------------------
        for (int l = tid; l < L; l += kRefThreads) {
            const MtRefineLine ln = p.lines[l];
            line_constants<NT, GENERAL>(cfg, cal, ln, xs[ln.comp], __ldg(p.line_window + l), l, L, lc0, lc1, lc2, lc3, lck);
        }
        __syncthreads();

        // ---- B1: channels -> model, theta derivatives, residual ---------------------------
        constexpr int kAcc = 1 + NT + NT * NT;  // loss, D^T r, D^T D
        f32x2 acc[kAcc > 0 ? kAcc : 1];
#pragma unroll
        for (int i = 0; i < kAcc; ++i) acc[i] = 0ull;
        int id_k[NT > 0 ? NT : 1];
#pragma unroll
        for (int k = 0; k < NT; ++k) id_k[k] = cfg.theta_id[k];
        const f32x2 off2 = bcast2(off), slope2 = bcast2(slope), quad2 = bcast2(cal.quad);
        // Channel pairs are visited in the order of pair_info: sorted by the number of lines that cover them, so that
        // the 32 pairs a warp evaluates together run the same number of line iterations (in channel order a warp spent
        // two thirds of its issue slots on lanes that had run out of lines: ncu r01_refine_v3, 78 thread instructions per
        // (line, pair) against 16 in the loop body).
        // the counts of the pair are requested before its lines are evaluated and used after (in visiting order the
        // loads are scattered: issued where they are used they stalled every pair for an L2 round trip)
        const int n_pairs = ld >> 1;
        for (int i = tid; i < n_pairs; i += kRefThreads) {
            const uint32_t info = __ldg(p.pair_info + i);
            float y0, y1;
            {
                const int c = 2 * (int)(info & 0xfffu);
                y0 = ld_spec<T>(row + c) - (brow ? __ldg(brow + c) : 0.f);
                y1 = c + 1 < n_ch ? ld_spec<T>(row + c + 1) - (brow ? __ldg(brow + c + 1) : 0.f) : 0.f;
            }
            const int c0 = 2 * (int)(info & 0xfffu);
            const bool has1 = c0 + 1 < n_ch;
            const float cg0 = (float)(cfg.ch0 + c0);
            const f32x2 cg = pack2(cg0, cg0 + 1.f);
            const int l_begin = (int)((info >> 12) & 0x3ffu);
            const int l_end = l_begin + (int)(info >> 22);
            f32x2 m, d[NT > 0 ? NT : 1];
            eval_pair<NT, GENERAL>(id_k, cal, cg, l_begin, l_end, L, lc0, lc1, lc2, lck, m, d);
            if (GENERAL && p.escape_bins > 0) {
                // escape peaks: the residual needs the model of channel c + bins as well; store the raw model and
                // derivative rows here, the shift passes below form the residual and the sums
                if (!has1) {
                    float lo, hi;
                    unpack2(m, lo, hi);
                    m = pack2(lo, 0.f);
#pragma unroll
                    for (int k = 0; k < NT; ++k) {
                        unpack2(d[k], lo, hi);
                        d[k] = pack2(lo, 0.f);
                    }
                }
                *reinterpret_cast<f32x2 *>(r_s + c0) = m;
#pragma unroll
                for (int k = 0; k < NT; ++k) *reinterpret_cast<f32x2 *>(d_s + k * ld + c0) = d[k];
                continue;
            }
            f32x2 res = add2(m, pack2(-y0, -y1));
            if (GENERAL && p.chan_w != nullptr) {
                // (only in the GENERAL instantiations, which the host selects for masked fits: the lean kernel of the
                // Gaussian-only benchmark path keeps its register budget)
                // channels outside `indices` leave the objective: their residual and derivative rows are zero, and with
                // them every sum the step is built from (loss, D^T r, D^T D here; B^T r, D^T B in the component pass)
                const f32x2 w2 = pack2(__ldg(p.chan_w + c0), has1 ? __ldg(p.chan_w + c0 + 1) : 0.f);
                res = mul2(res, w2);
#pragma unroll
                for (int k = 0; k < NT; ++k) d[k] = mul2(d[k], w2);
            }
            if (!has1) {
                // phantom channel of an odd window: contributes nothing anywhere
                float lo, hi;
                unpack2(res, lo, hi);
                res = pack2(lo, 0.f);
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    unpack2(d[k], lo, hi);
                    d[k] = pack2(lo, 0.f);
                }
            }
            *reinterpret_cast<f32x2 *>(r_s + c0) = res;
            acc[0] = fma2(res, res, acc[0]);
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                *reinterpret_cast<f32x2 *>(d_s + k * ld + c0) = d[k];
                acc[1 + k] = fma2(d[k], res, acc[1 + k]);
#pragma unroll
                for (int k2 = 0; k2 < NT; ++k2) acc[1 + NT + k * NT + k2] = fma2(d[k], d[k2], acc[1 + NT + k * NT + k2]);
            }
        }
        if (GENERAL && p.escape_bins > 0) {
            // ---- escape peaks (map.py:127-150): model'(c) = model(c) + f * model(c + bins) for c + bins < n_ch ------
            // In place, 128 channels at a time in ASCENDING order: a thread reads its channel and the shifted one, the
            // CTA synchronises, the thread writes -- the shifted source lies in the same block (hence the barrier) or in a
            // later one that is still untouched.  Residual, channel weights and the sums of the step are formed here.
            const int bins = p.escape_bins;
            const float f = p.escape_factor;
            __syncthreads();
            for (int base = 0; base < n_ch; base += kRefThreads) {
                const int c = base + tid;
                const bool in = c < n_ch;
                float mv = 0.f, dv[NT > 0 ? NT : 1];
#pragma unroll
                for (int k = 0; k < NT; ++k) dv[k] = 0.f;
                if (in) {
                    const bool sh = c + bins < n_ch;
                    mv = r_s[c] + (sh ? f * r_s[c + bins] : 0.f);
#pragma unroll
                    for (int k = 0; k < NT; ++k) dv[k] = d_s[k * ld + c] + (sh ? f * d_s[k * ld + c + bins] : 0.f);
                    const float y = ld_spec<T>(row + c) - (brow ? __ldg(brow + c) : 0.f);
                    const float w = p.chan_w != nullptr ? __ldg(p.chan_w + c) : 1.f;
                    mv = (mv - y) * w;
#pragma unroll
                    for (int k = 0; k < NT; ++k) dv[k] *= w;
                }
                __syncthreads();
                if (in) {
                    r_s[c] = mv;
                    const f32x2 r2 = pack2(mv, 0.f);
                    acc[0] = fma2(r2, r2, acc[0]);
#pragma unroll
                    for (int k = 0; k < NT; ++k) {
                        d_s[k * ld + c] = dv[k];
                        const f32x2 dk = pack2(dv[k], 0.f);
                        acc[1 + k] = fma2(dk, r2, acc[1 + k]);
#pragma unroll
                        for (int k2 = 0; k2 < NT; ++k2)
                            acc[1 + NT + k * NT + k2] = fma2(dk, pack2(dv[k2], 0.f), acc[1 + NT + k * NT + k2]);
                    }
                }
            }
            // The component pass multiplies the rows with the UNSHIFTED line shapes b(c); sum_c b'(c) r(c) equals
            // sum_c b(c) (r(c) + f * r(c - bins)), so the rows are folded once more, in DESCENDING order (the source
            // c - bins lies in the same block or in a lower one, processed later).
            __syncthreads();
            for (int base = ((n_ch - 1) / kRefThreads) * kRefThreads; base >= 0; base -= kRefThreads) {
                const int c = base + tid;
                const bool in = c < n_ch && c >= bins;
                float rv = 0.f, dv[NT > 0 ? NT : 1];
#pragma unroll
                for (int k = 0; k < NT; ++k) dv[k] = 0.f;
                if (in) {
                    rv = r_s[c] + f * r_s[c - bins];
#pragma unroll
                    for (int k = 0; k < NT; ++k) dv[k] = d_s[k * ld + c] + f * d_s[k * ld + c - bins];
                }
                __syncthreads();
                if (in) {
                    r_s[c] = rv;
#pragma unroll
                    for (int k = 0; k < NT; ++k) d_s[k * ld + c] = dv[k];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kAcc; ++i) {
            const float v = warp_sum(hsum2(acc[i]));
            if (lane == 0) {
                // scatter into the fixed [1 + kMaxTheta + kMaxTheta^2] layout the solve reads
                const int slot = i == 0 ? 0 : (i <= NT ? i : 1 + kMaxTheta + ((i - 1 - NT) / (NT > 0 ? NT : 1)) * kMaxTheta +
                                                                 (i - 1 - NT) % (NT > 0 ? NT : 1));
                part[warp][slot] = v;
            }
        }
        __syncthreads();

        // ---- B2: components -> B^T r and D^T B over the line windows ------------------------
        // window bounds are rounded outwards to even channels, so a lane owns the aligned pair (c, c + 1): 64-bit
        // shared-memory loads of r and d, packed arithmetic; r and d of the phantom channel are zero
        for (int slot = warp; slot < E; slot += kRefWarps) {
            const int e = __ldg(p.comp_order + slot);
            f32x2 g = 0ull, h[NT > 0 ? NT : 1];
#pragma unroll
            for (int k = 0; k < NT; ++k) h[k] = 0ull;
            if (p.b2_items != nullptr) {
                // flattened: the (line, channel pair) items of the component, 32 per warp iteration whatever the
                // lines' window lengths are (per-line window loops left a quarter of the lanes idle and spent more
                // instructions on loop set-up than on the items: ncu r02_refine_v4)
                const int i_end = __ldg(p.b2_start + e + 1);
                int i = __ldg(p.b2_start + e) + lane;
                uint32_t item_next = i < i_end ? __ldg(p.b2_items + i) : 0x3ffu;
                for (; i < i_end; i += 32) {
                    const uint32_t item = item_next;  // fetched an iteration ago: the table comes from L2 / L1
                    if (i + 32 < i_end) item_next = __ldg(p.b2_items + i + 32);
                    const int l = (int)(item & 0x3ffu);
                    if (l == 0x3ff) continue;
                    const int c = 2 * (int)(item >> 10);
                    const float4 a0 = lc0[l];
                    const float cg0 = (float)(cfg.ch0 + c);
                    const f32x2 cg = pack2(cg0, cg0 + 1.f);
                    const f32x2 dl = add2(fma2(fma2(quad2, cg, slope2), cg, off2), bcast2(-a0.x));
                    float e0, e1;
                    unpack2(mul2(mul2(dl, dl), bcast2(a0.y)), e0, e1);
                    f32x2 b = mul2(pack2(ex2_approx(e0), ex2_approx(e1)), bcast2(lc1[l].w));
                    if (GENERAL) {
                        const float kstep = lc3[l].x;
                        if (kstep != 0.f) {
                            float r0, r1;
                            erfc2(mul2(dl, bcast2(lc2[l].z)), r0, r1);
                            b = fma2(pack2(r0, r1), bcast2(kstep), b);
                        }
                    }
                    const float *rp = r_s + c;
                    g = fma2(b, *reinterpret_cast<const f32x2 *>(rp), g);
#pragma unroll
                    for (int k = 0; k < NT; ++k) h[k] = fma2(b, *reinterpret_cast<const f32x2 *>(rp + (k + 1) * ld), h[k]);
                }
            }
            const int j_end = p.b2_items != nullptr ? 0 : __ldg(p.comp_start + e + 1);
            for (int j = __ldg(p.comp_start + e); j < j_end; ++j) {
                const int l = __ldg(p.comp_lines + j);
                const float4 a0 = lc0[l];
                const uint32_t win = __float_as_uint(a0.w);
                const f32x2 unit2 = bcast2(lc1[l].w), ay2 = bcast2(a0.y), me2 = bcast2(-a0.x);
                float kstep = 0.f, zs = 0.f;
                if (GENERAL) {
                    kstep = lc3[l].x;
                    zs = lc2[l].z;
                }
                const int c_end = min(((int)(win >> 16) + 1) & ~1, ld);
                int c = ((int)(win & 0xffffu) & ~1) + 2 * lane;
                const float cg0 = (float)(cfg.ch0 + c);
                f32x2 cg = pack2(cg0, cg0 + 1.f);
                const float *rp = r_s + c;
                for (; c < c_end; c += 64, cg = add2(cg, bcast2(64.f)), rp += 64) {
                    const f32x2 dl = add2(fma2(fma2(quad2, cg, slope2), cg, off2), me2);
                    float e0, e1;
                    unpack2(mul2(mul2(dl, dl), ay2), e0, e1);
                    f32x2 b = mul2(pack2(ex2_approx(e0), ex2_approx(e1)), unit2);
                    if (GENERAL && kstep != 0.f) {
                        float r0, r1;
                        erfc2(mul2(dl, bcast2(zs)), r0, r1);
                        b = fma2(pack2(r0, r1), bcast2(kstep), b);
                    }
                    g = fma2(b, *reinterpret_cast<const f32x2 *>(rp), g);
#pragma unroll
                    for (int k = 0; k < NT; ++k) h[k] = fma2(b, *reinterpret_cast<const f32x2 *>(rp + (k + 1) * ld), h[k]);
                }
            }
            const float gs = warp_sum(hsum2(g));
            float hs[NT > 0 ? NT : 1];
#pragma unroll
            for (int k = 0; k < NT; ++k) hs[k] = warp_sum(hsum2(h[k]));
            if (lane == 0) {
                gx[e] = gs;
#pragma unroll
                for (int k = 0; k < NT; ++k) hthx[k][e] = hs[k];
            }
        }
        __syncthreads();

        // ---- accept / reject the previous step -----------------------------------------------
        // every thread forms the same decision from the same shared values: no hand-off, no barrier
        float loss = 0.f;
#pragma unroll
        for (int w = 0; w < kRefWarps; ++w) loss += part[w][0];
        // a step is rejected only when the loss rises by more than its evaluation noise (2-ulp exponentials,
        // fp32 sums: ~1e-6 relative); smaller rises are what the last Newton steps look like in fp32 and
        // rejecting them costs ten futile step halvings per pixel
        const bool rejected = loss > ctl[1] * (1.f + 2e-5f) && iter > 0;
        if (tid == 0) ctl[0] = loss;
        if (rejected) {
            // go back and retry with half the step (costs one evaluation)
            const float alpha = ctl[2] * 0.5f;
            for (int e = tid; e < E; e += kRefThreads) xs[e] = fmaxf(x_prev[e] + alpha * dx[e], 0.f);
            if (tid < NT) theta[tid] = fminf(fmaxf(theta_prev[tid] + alpha * dtheta[tid], cfg.theta_lo[tid]), cfg.theta_hi[tid]);
            __syncthreads();
            if (tid == 0) ctl[2] = alpha;
            __syncthreads();
            if (alpha < 1e-3f || iter >= cfg.max_iter + 8) {
                // give up on this direction: restore the last accepted state and stop
                for (int e = tid; e < E; e += kRefThreads) xs[e] = x_prev[e];
                if (tid < NT) theta[tid] = theta_prev[tid];
                if (tid == 0) ctl[0] = ctl[1];
                __syncthreads();
                break;
            }
            continue;
        }

        // ---- C: damped Gauss-Newton step through the Schur complement on theta ----------------
        bool clamped_here = false;  // does this thread see a component held at zero by its gradient?
        for (int i = tid; i < E * (NT + 1); i += kRefThreads) {
            const int e = i / (NT + 1), k = i % (NT + 1);
            const float *gi = p.g0inv + (long long)e * E;
            float s = 0.f;
            if (k == NT) {
                for (int j = 0; j < E; ++j) s += __ldg(gi + j) * gx[j];
                w_s[e] = s;
                clamped_here = clamped_here || (xs[e] <= 0.f && gx[e] > 0.f);
            } else {
                for (int j = 0; j < E; ++j) s += __ldg(gi + j) * hthx[k][j];
                t1[e][k] = s;
            }
        }
        // (the barrier this loop needs anyway also tells every thread whether any component is clamped)
        const bool any_clamped = __syncthreads_or(clamped_here) != 0;
        // ---- active set: components at zero whose gradient keeps them there -----------------------------------------
        // w_s and t1 were formed with the inverse of the FULL Gram matrix, i.e. they describe a step in which a clamped
        // amplitude is free to go negative; projecting that step afterwards biases every other component and the
        // iteration stalls where the least-squares start left it (pixels without one of the fitted elements).  The
        // Newton step has to be taken on the free set F alone, and the inverse of G_FF follows from H = G^-1 by
        // eliminating the active rows:  (G_FF)^-1 = H_FF - H_FA (H_AA)^-1 H_AF,  i.e.  V <- V - H[:, A] (H_AA)^-1 V[A]
        // for the NT + 1 vectors V = (t1, w_s).  A = { e : x_e = 0 and (B^T r)_e > 0 }; the scratch is the residual /
        // derivative rows, dead until the next evaluation.  Pixels with an empty A (the common case) skip all of it.
        if (any_clamped) {
            __shared__ int s_na;
            __shared__ unsigned char s_act[kMaxComp], s_in[kMaxComp];
            constexpr int NV = NT + 1;
            if (warp == 0) {
                int base = 0;
                for (int e0 = 0; e0 < E; e0 += 32) {
                    const int e = e0 + lane;
                    const bool act = e < E && xs[e] <= 0.f && gx[e] > 0.f;
                    const unsigned m = __ballot_sync(0xffffffffu, act);
                    const int pos = base + __popc(m & ((1u << lane) - 1u));
                    if (e < E) s_in[e] = act && pos < cap;
                    if (act && pos < cap) s_act[pos] = (unsigned char)e;
                    base += __popc(m);
                }
                if (lane == 0) s_na = min(base, cap);
            }
            __syncthreads();
            const int na = s_na;
            if (na > 0) {
                float *M = r_s, *R = r_s + na * na;  // M = H_AA [na][na], R = V[A] [na][NV]
                for (int i = tid; i < na * na; i += kRefThreads)
                    M[i] = __ldg(p.g0inv + (long long)s_act[i / na] * E + s_act[i % na]);
                for (int i = tid; i < na * NV; i += kRefThreads) {
                    const int e = s_act[i / NV], k = i % NV;
                    R[i] = k == NT ? w_s[e] : t1[e][k < NT ? k : 0];
                }
                // elimination without pivoting (H_AA is symmetric positive definite), thread q owns row q
                for (int c = 0; c < na; ++c) {
                    __syncthreads();
                    if (tid > c && tid < na) {
                        const float f = M[tid * na + c] / M[c * na + c];
                        for (int cc = c + 1; cc < na; ++cc) M[tid * na + cc] = fmaf(-f, M[c * na + cc], M[tid * na + cc]);
                        for (int k = 0; k < NV; ++k) R[tid * NV + k] = fmaf(-f, R[c * NV + k], R[tid * NV + k]);
                    }
                }
                for (int c = na - 1; c >= 0; --c) {
                    __syncthreads();
                    if (tid < NV) R[c * NV + tid] = R[c * NV + tid] / M[c * na + c];
                    __syncthreads();
                    if (tid < c)
                        for (int k = 0; k < NV; ++k) R[tid * NV + k] = fmaf(-M[tid * na + c], R[c * NV + k], R[tid * NV + k]);
                }
                __syncthreads();
                for (int i = tid; i < E * NV; i += kRefThreads) {
                    const int e = i / NV, k = i % NV;
                    float corr = 0.f;
                    for (int j = 0; j < na; ++j) corr = fmaf(__ldg(p.g0inv + (long long)e * E + s_act[j]), R[j * NV + k], corr);
                    if (!isfinite(corr)) corr = 0.f;  // (a non-positive pivot in fp32: leave the entry as it was)
                    const float v = s_in[e] ? 0.f : (k == NT ? w_s[e] : t1[e][k < NT ? k : 0]) - corr;
                    if (k == NT)
                        w_s[e] = v;
                    else
                        t1[e][k < NT ? k : 0] = v;
                }
                __syncthreads();
            }
        }
        // Schur complement of the theta block: the E-long dot products across warp 0, the NT x NT solve on lane 0
        __shared__ float schur[kMaxTheta + kMaxTheta * kMaxTheta];
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                float g = 0.f;
                for (int j = lane; j < E; j += 32) g = fmaf(hthx[k][j], w_s[j], g);
                g = warp_sum(g);
                if (lane == 0) schur[k] = g;
#pragma unroll
                for (int k2 = 0; k2 < NT; ++k2) {
                    float h = 0.f;
                    for (int j = lane; j < E; j += 32) h = fmaf(hthx[k][j], t1[j][k2], h);
                    h = warp_sum(h);
                    if (lane == 0) schur[kMaxTheta + k * kMaxTheta + k2] = h;
                }
            }
            __syncwarp();
        }
        if (tid == 0) {
            float S[NT > 0 ? NT : 1][NT > 0 ? NT : 1], rhs[NT > 0 ? NT : 1], sol[NT > 0 ? NT : 1];
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                float g = -schur[k];
                for (int w = 0; w < kRefWarps; ++w) g += part[w][1 + k];
                rhs[k] = -g;
#pragma unroll
                for (int k2 = 0; k2 < NT; ++k2) {
                    float h = -schur[kMaxTheta + k * kMaxTheta + k2];
                    for (int w = 0; w < kRefWarps; ++w) h += part[w][1 + kMaxTheta + k * kMaxTheta + k2];
                    S[k][k2] = h;
                }
            }
#pragma unroll
            for (int k = 0; k < NT; ++k) S[k][k] = S[k][k] * 1.0001f + 1e-30f;
            // S is the Schur complement of a positive semi-definite Gauss-Newton matrix (plus the damping above):
            // symmetric positive definite, so elimination needs no pivoting; fully unrolled, everything in registers
            bool ok = true;
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const float pv = S[c][c];
                ok = ok && pv > 0.f && isfinite(pv);
                const float inv = 1.f / pv;
#pragma unroll
                for (int q = c + 1; q < NT; ++q) {
                    const float f = S[q][c] * inv;
#pragma unroll
                    for (int cc = c; cc < NT; ++cc) S[q][cc] -= f * S[c][cc];
                    rhs[q] -= f * rhs[c];
                }
            }
#pragma unroll
            for (int c = NT - 1; c >= 0; --c) {
                float acc_s = rhs[c];
#pragma unroll
                for (int cc = c + 1; cc < NT; ++cc) acc_s -= S[c][cc] * sol[cc];
                sol[c] = ok ? acc_s / S[c][c] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                theta_prev[k] = theta[k];
                const float step = isfinite(sol[k]) ? sol[k] : 0.f;
                const float nt = fminf(fmaxf(theta[k] + step, cfg.theta_lo[k]), cfg.theta_hi[k]);
                dtheta[k] = nt - theta[k];
                theta[k] = nt;
            }
            ctl[1] = ctl[0];
            ctl[2] = 1.f;
        }
        __syncthreads();
        float change = 0.f;
        for (int e = tid; e < E; e += kRefThreads) {
            float step = -w_s[e];
            for (int k = 0; k < NT; ++k) step -= t1[e][k] * dtheta[k];
            x_prev[e] = xs[e];
            const float nx = fmaxf(xs[e] + step, 0.f);
            dx[e] = nx - xs[e];
            xs[e] = nx;
            change = fmaxf(change, fabsf(dx[e]));
        }
        // convergence: relative amplitude change and theta steps below tol
        float xmax = 0.f;
        for (int e = tid; e < E; e += kRefThreads) xmax = fmaxf(xmax, fabsf(xs[e]));
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            change = fmaxf(change, __shfl_xor_sync(0xffffffffu, change, s));
            xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, s));
        }
        __shared__ float cmax[kRefWarps], xm[kRefWarps];
        if (lane == 0) {
            cmax[warp] = change;
            xm[warp] = xmax;
        }
        __syncthreads();
        float cm = 0.f, xmm = 0.f;
        for (int w = 0; w < kRefWarps; ++w) {
            cm = fmaxf(cm, cmax[w]);
            xmm = fmaxf(xmm, xm[w]);
        }
        bool small = cm <= cfg.tol * fmaxf(xmm, 1e-30f);
        for (int k = 0; k < NT; ++k) small = small && fabsf(dtheta[k]) <= cfg.tol * fmaxf(fabsf(cfg.theta_hi[k] - cfg.theta_lo[k]), 1e-30f);
        __syncthreads();
        if (iter >= cfg.max_iter) {
            // the state was just stepped; report the last evaluated state instead (its loss is known)
            for (int e = tid; e < E; e += kRefThreads) xs[e] = x_prev[e];
            if (tid < NT) theta[tid] = theta_prev[tid];
            __syncthreads();
            break;
        }
        if (small && iter > 0) {
            // converged: one more evaluation would only confirm; keep the stepped state, loss of the
            // previous evaluation is an upper bound to within tol
            break;
        }
    }

    for (int e = tid; e < E; e += kRefThreads) p.x[(long long)e * p.ld_x + px] = xs[e];
    if (tid < NT && p.theta_out) p.theta_out[(long long)tid * p.ld_theta + px] = theta[tid];
    if (tid == 0) {
        if (p.loss_out) p.loss_out[px] = ctl[0];
        if (p.iters_out) p.iters_out[px] = iter;
    }
}

// ---- model and d model / d theta of one spectrum (the Jacobian of fit_spec(tune_params=True)) ---------------------
template <int NT>
__global__ void __launch_bounds__(kRefThreads)
jacobian_kernel(const MtRefineConfig cfg, const MtRefineLine *__restrict__ lines, const uint32_t *__restrict__ chan_lines,
                const float *__restrict__ x, float *__restrict__ model, float *__restrict__ dtheta, long long ld_out) {
    extern __shared__ float4 dyn4[];
    const int L = cfg.n_lines, n_ch = cfg.n_ch;
    float4 *lc0 = dyn4, *lc1 = dyn4 + L, *lc2 = dyn4 + 2 * L, *lc3 = dyn4 + 3 * L, *lck = dyn4 + 4 * L;
    const Calib cal = calib_of(cfg, cfg.theta0, NT);
    for (int l = threadIdx.x; l < L; l += kRefThreads) {
        const MtRefineLine ln = lines[l];
        line_constants<NT, true>(cfg, cal, ln, fmaxf(__ldg(x + ln.comp), 0.f), 0u, l, L, lc0, lc1, lc2, lc3, lck);
    }
    __syncthreads();
    int id_k[NT > 0 ? NT : 1];
#pragma unroll
    for (int k = 0; k < NT; ++k) id_k[k] = cfg.theta_id[k];
    const int q = blockIdx.x * kRefThreads + threadIdx.x;
    const int c0 = 2 * q;
    if (c0 >= n_ch) return;
    const bool has1 = c0 + 1 < n_ch;
    const float cg0 = (float)(cfg.ch0 + c0);
    const uint32_t ra = __ldg(chan_lines + c0), rb = has1 ? __ldg(chan_lines + c0 + 1) : ra;
    const int l_begin = min((int)(ra & 0xffffu), (int)(rb & 0xffffu));
    const int l_end = max((int)(ra >> 16), (int)(rb >> 16));
    f32x2 m, d[NT > 0 ? NT : 1];
    eval_pair<NT, true>(id_k, cal, pack2(cg0, cg0 + 1.f), l_begin, l_end, L, lc0, lc1, lc2, lck, m, d);
    float lo, hi;
    unpack2(m, lo, hi);
    model[c0] = lo;
    if (has1) model[c0 + 1] = hi;
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        unpack2(d[k], lo, hi);
        dtheta[k * ld_out + c0] = lo;
        if (has1) dtheta[k * ld_out + c0 + 1] = hi;
    }
}

// ---- device-resident copies of the host tables, keyed by content -------------------------------------------
struct TableEntry {
    int device = -1;
    uint64_t hash = 0;
    std::vector<char> host;  // full comparison on a hash match
    char *dev = nullptr;
    uint64_t stamp = 0;
};
constexpr int kTableSlots = 16;
TableEntry g_tables[kTableSlots];
std::mutex g_tables_mutex;
uint64_t g_tables_clock = 0;

int cached_tables(const std::vector<char> &packed, const char **out) {
    uint64_t h = 1469598103934665603ull;  // FNV-1a over 8-byte words (the buffer is padded to 256-byte pieces)
    const uint64_t *w = reinterpret_cast<const uint64_t *>(packed.data());
    for (size_t i = 0; i < packed.size() / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    int dev = 0;
    MT_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    int victim = 0;
    for (int i = 0; i < kTableSlots; ++i) {
        TableEntry &t = g_tables[i];
        if (t.dev && t.device == dev && t.hash == h && t.host == packed) {
            t.stamp = ++g_tables_clock;
            *out = t.dev;
            return MT_OK;
        }
        if (g_tables[i].stamp < g_tables[victim].stamp) victim = i;
    }
    TableEntry &t = g_tables[victim];
    if (t.dev) {
        // a kernel of an earlier call may still be reading the evicted tables
        int cur = dev;
        if (t.device != cur) cudaSetDevice(t.device);
        cudaDeviceSynchronize();
        cudaFree(t.dev);
        if (t.device != cur) cudaSetDevice(cur);
        t.dev = nullptr;
    }
    MT_CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&t.dev), packed.size()));
    MT_CUDA_TRY(cudaMemcpy(t.dev, packed.data(), packed.size(), cudaMemcpyHostToDevice));  // synchronous: visible to every stream
    t.device = dev;
    t.hash = h;
    t.host = packed;
    t.stamp = ++g_tables_clock;
    *out = t.dev;
    return MT_OK;
}

}  // namespace

extern "C" int mt_refine(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, const float *bkg_dev,
                         int64_t ld_bkg, const float *chan_weight_dev, int32_t escape_bins, float escape_factor,
                         const MtRefineConfig *cfg_host, const MtRefineLine *lines_host,
                         const uint32_t *chan_lines_host, const int32_t *comp_start_host,
                         const int32_t *comp_lines_host, const uint32_t *line_window_host,
                         const int32_t *comp_order_host, const float *g0inv_dev,
                         float *x_dev, int64_t ld_x, float *theta_dev, int64_t ld_theta, float *loss_dev, int32_t *iters_dev,
                         void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && cfg_host && lines_host && chan_lines_host && comp_start_host && comp_lines_host &&
                   line_window_host && comp_order_host && g0inv_dev && x_dev,
               "mt_refine: null pointer");
    const MtRefineConfig &cfg = *cfg_host;
    MT_REQUIRE(cfg.n_ch > 0 && cfg.n_ch <= 8192 && cfg.ch0 >= 0 && ld_spec >= (int64_t)cfg.ch0 + cfg.n_ch,
               "mt_refine: channel window");
    MT_REQUIRE(cfg.n_comp >= 1 && cfg.n_comp <= kMaxComp, "mt_refine: n_comp=%d (max %d)", cfg.n_comp, kMaxComp);
    MT_REQUIRE(cfg.n_lines >= 1 && cfg.n_lines <= kMaxLines, "mt_refine: n_lines=%d (max %d)", cfg.n_lines, kMaxLines);
    MT_REQUIRE(cfg.n_theta >= 0 && cfg.n_theta <= kMaxTheta, "mt_refine: n_theta=%d (at most %d per pixel)", cfg.n_theta,
               kMaxTheta);
    MT_REQUIRE(n_px >= 0 && ld_x >= n_px && (!bkg_dev || ld_bkg >= cfg.n_ch), "mt_refine: sizes");
    MT_REQUIRE(escape_bins >= 0 && escape_bins < cfg.n_ch && escape_factor >= 0.f, "mt_refine: escape peak (bins %d, factor %g)",
               escape_bins, (double)escape_factor);
    MT_REQUIRE(!theta_dev || cfg.n_theta == 0 || ld_theta >= n_px, "mt_refine: ld_theta=%lld < n_px", (long long)ld_theta);
    for (int l = 0; l < cfg.n_lines; ++l)
        MT_REQUIRE(lines_host[l].comp >= 0 && lines_host[l].comp < cfg.n_comp && lines_host[l].kind >= MT_LINE_ELEMENT &&
                       lines_host[l].kind <= MT_LINE_COMPTON && lines_host[l].step >= 0.f,
                   "mt_refine: line %d (component / kind / step)", l);
    for (int k = 0; k < cfg.n_theta; ++k)
        MT_REQUIRE(cfg.theta_id[k] >= 0 && cfg.theta_id[k] < MT_N_THETA_IDS, "mt_refine: theta_id[%d]=%d", k, cfg.theta_id[k]);
    {
        bool seen[kMaxComp] = {false};
        for (int e = 0; e < cfg.n_comp; ++e) {
            const int c = comp_order_host[e];
            MT_REQUIRE(c >= 0 && c < cfg.n_comp && !seen[c], "mt_refine: comp_order is not a permutation (entry %d)", e);
            seen[c] = true;
        }
    }
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    // the small tables, packed into one host buffer
    const size_t b_lines = sizeof(MtRefineLine) * cfg.n_lines, b_chan = sizeof(uint32_t) * cfg.n_ch,
                 b_cs = sizeof(int32_t) * (cfg.n_comp + 1), b_cl = sizeof(int32_t) * cfg.n_lines,
                 b_lw = sizeof(uint32_t) * cfg.n_lines, b_co = sizeof(int32_t) * cfg.n_comp;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // channel pairs sorted by the number of lines covering them (ties in channel order): the visiting order of phase B1
    const int n_pairs = (cfg.n_ch + 1) / 2;
    std::vector<uint32_t> pair_info(n_pairs);
    {
        std::vector<int> count(n_pairs), first(n_pairs), order(n_pairs);
        for (int q = 0; q < n_pairs; ++q) {
            const uint32_t ra = chan_lines_host[2 * q], rb = 2 * q + 1 < cfg.n_ch ? chan_lines_host[2 * q + 1] : ra;
            // union of the two channels' line ranges; a channel without lines (empty range, encoded 0 | 0) must not
            // stretch its neighbour's range down to line 0
            const int la = (int)(ra & 0xffffu), ha = (int)(ra >> 16), lb = (int)(rb & 0xffffu), hb = (int)(rb >> 16);
            const int lo = ha <= la ? lb : (hb <= lb ? la : std::min(la, lb));
            const int hi = ha <= la ? hb : (hb <= lb ? ha : std::max(ha, hb));
            MT_REQUIRE(lo >= 0 && hi <= cfg.n_lines && hi - lo < 1024 && lo < 1024, "mt_refine: chan_lines[%d] out of range", 2 * q);
            first[q] = lo;
            count[q] = hi > lo ? hi - lo : 0;
            order[q] = q;
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return count[a] > count[b]; });
        for (int i = 0; i < n_pairs; ++i)
            pair_info[i] = (uint32_t)order[i] | ((uint32_t)first[order[i]] << 12) | ((uint32_t)count[order[i]] << 22);
    }
    const size_t b_pi = sizeof(uint32_t) * n_pairs;
    // component-major pass: every component's (line, channel pair) items in one list, padded to whole warps
    std::vector<uint32_t> b2_items;
    std::vector<int32_t> b2_start(cfg.n_comp + 1, 0);
    bool flat_b2 = cfg.n_lines < 0x3ff;
    for (int e = 0; e < cfg.n_comp && flat_b2; ++e) {
        b2_start[e] = (int32_t)b2_items.size();
        for (int j = comp_start_host[e]; j < comp_start_host[e + 1]; ++j) {
            const int l = comp_lines_host[j];
            MT_REQUIRE(l >= 0 && l < cfg.n_lines, "mt_refine: comp_lines[%d]=%d", j, l);
            const uint32_t win = line_window_host[l];
            const int q0 = (int)(win & 0xffffu) / 2, q1 = std::min(((int)(win >> 16) + 1) / 2, n_pairs);
            for (int q = q0; q < q1; ++q) b2_items.push_back((uint32_t)l | ((uint32_t)q << 10));
        }
        while (b2_items.size() % 32) b2_items.push_back(0x3ffu);
        flat_b2 = b2_items.size() <= (1u << 18);  // 1 MB of table at most; beyond that the per-line loops serve
    }
    b2_start[cfg.n_comp] = (int32_t)b2_items.size();
    if (!flat_b2) b2_items.assign(1, 0x3ffu);
    const size_t b_it = sizeof(uint32_t) * b2_items.size(), b_is = sizeof(int32_t) * b2_start.size();
    const size_t total = up(b_lines) + up(b_chan) + up(b_cs) + up(b_cl) + up(b_lw) + up(b_co) + up(b_pi) + up(b_it) + up(b_is);
    std::vector<char> packed(total);
    size_t cursor = 0;
    RefParams p{};
    p.cfg = cfg;
    size_t o_lines, o_chan, o_cs, o_cl, o_lw, o_co;
    auto stage = [&](const void *src, size_t bytes) -> size_t {
        memcpy(packed.data() + cursor, src, bytes);
        const size_t at = cursor;
        cursor += up(bytes);
        return at;
    };
    o_lines = stage(lines_host, b_lines);
    o_chan = stage(chan_lines_host, b_chan);
    o_cs = stage(comp_start_host, b_cs);
    o_cl = stage(comp_lines_host, b_cl);
    o_lw = stage(line_window_host, b_lw);
    o_co = stage(comp_order_host, b_co);
    const size_t o_pi = stage(pair_info.data(), b_pi);
    const size_t o_it = stage(b2_items.data(), b_it), o_is = stage(b2_start.data(), b_is);
    // The tables are the same for every slab of a map.  Uploading them per call would put a host-to-device copy
    // in front of every launch, and that copy queues on the DMA engine BEHIND the next slab of the volume that
    // a streaming caller is already sending -- the kernel would wait for it.  So device copies are kept in a small
    // content-addressed cache; a miss uploads synchronously (once per plan), a hit costs one hash of ~10 KB.
    const char *tables = nullptr;
    if (int rc = cached_tables(packed, &tables)) return rc;
    p.lines = reinterpret_cast<const MtRefineLine *>(tables + o_lines);
    p.chan_lines = reinterpret_cast<const uint32_t *>(tables + o_chan);
    p.comp_start = reinterpret_cast<const int32_t *>(tables + o_cs);
    p.comp_lines = reinterpret_cast<const int32_t *>(tables + o_cl);
    p.line_window = reinterpret_cast<const uint32_t *>(tables + o_lw);
    p.comp_order = reinterpret_cast<const int32_t *>(tables + o_co);
    p.pair_info = reinterpret_cast<const uint32_t *>(tables + o_pi);
    p.b2_items = flat_b2 ? reinterpret_cast<const uint32_t *>(tables + o_it) : nullptr;
    p.b2_start = reinterpret_cast<const int32_t *>(tables + o_is);
    p.g0inv = g0inv_dev;
    p.bkg = bkg_dev;
    p.ld_bkg = ld_bkg;
    p.chan_w = chan_weight_dev;
    p.escape_bins = escape_bins > 0 && escape_factor > 0.f ? escape_bins : 0;
    p.escape_factor = escape_factor;
    p.spec = spec_dev;
    p.ld_spec = ld_spec;
    p.x = x_dev;
    p.ld_x = ld_x;
    p.theta_out = theta_dev;
    p.ld_theta = ld_theta;
    p.loss_out = loss_dev;
    p.iters_out = iters_dev;
    p.n_px = n_px;
    // GENERAL instantiations carry the step terms and the scatter-line parameters; plain Gaussian line lists with
    // axis / width parameters keep the lean kernel
    bool general = false;
    for (int k = 0; k < cfg.n_theta; ++k) general = general || cfg.theta_id[k] >= MT_THETA_COHERENT_SCT_ENERGY;
    for (int l = 0; l < cfg.n_lines; ++l) general = general || lines_host[l].step > 0.f;
    general = general || chan_weight_dev != nullptr;  // the channel weights live in the GENERAL instantiations
    general = general || escape_bins > 0;              // ... and so do the escape-peak passes
    const size_t smem = sizeof(float) * (size_t)((cfg.n_ch + 1) & ~1) * (1 + cfg.n_theta) +
                        sizeof(float4) * (size_t)cfg.n_lines * (general ? 4 + cfg.n_theta : 2);
    if (smem > 200 * 1024)
        return mt::fail(MT_ERR_UNSUPPORTED, "mt_refine: %zu bytes of shared memory needed (channels x free parameters)",
                        smem);
#define MT_LAUNCH_REFINE_NT(T, N, G)                                                                                     \
    do {                                                                                                                 \
        MT_CUDA_TRY(cudaFuncSetAttribute(refine_kernel<T, N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        refine_kernel<T, N, G><<<(unsigned)n_px, kRefThreads, smem, stream>>>(p);                                        \
    } while (0)
#define MT_LAUNCH_REFINE_G(T, G)                                \
    do {                                                        \
        switch (cfg.n_theta) {                                  \
            case 0: MT_LAUNCH_REFINE_NT(T, 0, G); break;        \
            case 1: MT_LAUNCH_REFINE_NT(T, 1, G); break;        \
            case 2: MT_LAUNCH_REFINE_NT(T, 2, G); break;        \
            case 3: MT_LAUNCH_REFINE_NT(T, 3, G); break;        \
            default: MT_LAUNCH_REFINE_NT(T, 4, G); break;       \
        }                                                       \
    } while (0)
#define MT_LAUNCH_REFINE(T)                                     \
    do {                                                        \
        if (general) MT_LAUNCH_REFINE_G(T, true);               \
        else MT_LAUNCH_REFINE_G(T, false);                      \
    } while (0)
    if (dtype == MT_DTYPE_F32) {
        MT_LAUNCH_REFINE(float);
    } else if (dtype == MT_DTYPE_F16) {
        MT_LAUNCH_REFINE(__half);
    } else {
        return mt::fail(MT_ERR_UNSUPPORTED, "mt_refine: dtype %d", dtype);
    }
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_model_jacobian(const MtRefineConfig *cfg_host, const MtRefineLine *lines_host, const uint32_t *chan_lines_host,
                                 const float *x_dev, float *model_dev, float *dtheta_dev, int64_t ld, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(cfg_host && lines_host && chan_lines_host && x_dev && model_dev, "mt_model_jacobian: null pointer");
    const MtRefineConfig &cfg = *cfg_host;
    MT_REQUIRE(cfg.n_ch > 0 && cfg.n_ch <= 8192 && cfg.ch0 >= 0, "mt_model_jacobian: channel window");
    MT_REQUIRE(cfg.n_lines >= 1 && cfg.n_lines <= 2048 && cfg.n_comp >= 1, "mt_model_jacobian: n_lines=%d n_comp=%d", cfg.n_lines,
               cfg.n_comp);
    MT_REQUIRE(cfg.n_theta >= 0 && cfg.n_theta <= MT_MAX_THETA && (cfg.n_theta == 0 || (dtheta_dev && ld >= cfg.n_ch)),
               "mt_model_jacobian: n_theta=%d ld=%lld", cfg.n_theta, (long long)ld);
    for (int l = 0; l < cfg.n_lines; ++l)
        MT_REQUIRE(lines_host[l].comp >= 0 && lines_host[l].comp < cfg.n_comp && lines_host[l].kind >= MT_LINE_ELEMENT &&
                       lines_host[l].kind <= MT_LINE_COMPTON && lines_host[l].step >= 0.f,
                   "mt_model_jacobian: line %d (component / kind / step)", l);
    for (int k = 0; k < cfg.n_theta; ++k)
        MT_REQUIRE(cfg.theta_id[k] >= 0 && cfg.theta_id[k] < MT_N_THETA_IDS, "mt_model_jacobian: theta_id[%d]=%d", k,
                   cfg.theta_id[k]);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t b_lines = sizeof(MtRefineLine) * cfg.n_lines, b_chan = sizeof(uint32_t) * cfg.n_ch;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    std::vector<char> packed(up(b_lines) + up(b_chan));
    memcpy(packed.data(), lines_host, b_lines);
    memcpy(packed.data() + up(b_lines), chan_lines_host, b_chan);
    const char *tables = nullptr;
    if (int rc = cached_tables(packed, &tables)) return rc;
    const MtRefineLine *lines = reinterpret_cast<const MtRefineLine *>(tables);
    const uint32_t *chan = reinterpret_cast<const uint32_t *>(tables + up(b_lines));
    const size_t smem = sizeof(float4) * (size_t)cfg.n_lines * (4 + cfg.n_theta);
    if (smem > 200 * 1024) return mt::fail(MT_ERR_UNSUPPORTED, "mt_model_jacobian: %zu bytes of shared memory needed", smem);
    const unsigned grid = (unsigned)(((cfg.n_ch + 1) / 2 + kRefThreads - 1) / kRefThreads);
#define MT_LAUNCH_JAC(N)                                                                                              \
    case N:                                                                                                           \
        MT_CUDA_TRY(cudaFuncSetAttribute(jacobian_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        jacobian_kernel<N><<<grid, kRefThreads, smem, stream>>>(cfg, lines, chan, x_dev, model_dev, dtheta_dev, ld);   \
        break;
    switch (cfg.n_theta) {
        MT_LAUNCH_JAC(0) MT_LAUNCH_JAC(1) MT_LAUNCH_JAC(2) MT_LAUNCH_JAC(3) MT_LAUNCH_JAC(4)
        MT_LAUNCH_JAC(5) MT_LAUNCH_JAC(6) MT_LAUNCH_JAC(7) MT_LAUNCH_JAC(8)
    }
#undef MT_LAUNCH_JAC
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
