// K2 -- SNIP background kernel.
//
// Replaces snip_bkg/snip_op (bkg.py:53-114): boxcar smooth -> log1p -> n_pass clipping passes
// bg[i] <- min((bg[lo_i] + bg[hi_i]) / 2, bg[i]) -> expm1, for a batch of spectra.  The
// reference runs ~10 ATen ops plus a host sync per pass over HBM-resident [P,C] tensors; here
// a CTA keeps FOUR spectra on chip for all passes:
//   * shared memory holds the log-background channel-major as float4 = the 4 pixels of the CTA,
//     so one LDS.128 fetches a neighbour for all four pixels with a single address;
//   * every thread owns channels {tid + T*k} and keeps their current values in registers (the
//     centre value never touches shared memory): 2 LDS.128 + 1 STS.128 per channel and pass;
//   * the gather indices depend only on global parameters, so they are precomputed on the
//     host bit-exactly; the words of the next pass are prefetched into registers
//     while the current pass computes (they come from L2).
// Bound: shared-memory bandwidth (12 bytes per channel, pass and pixel), not HBM.
#include <cuda_fp16.h>
#include <cstdlib>

#include "mt_common.cuh"
#include "mt_f32x2.cuh"
#include "mt_snip_dev.cuh"

namespace {

constexpr int kSnipThreads = 512;
constexpr int kSnipPix = 4;

using namespace mt_snip_dev;

// KCH channels per thread (n_ch <= kSnipThreads * KCH); 4 pixels per CTA.
// SCALED: the gather table holds shared-memory BYTE offsets (lo * 16 | hi * 16 << 16, mt_snip_scale_tables), its rows
// are padded to the thread grid and followed by a zero row, so the pass loop needs no shifts and no bounds tests.
template <typename T, int KCH, bool SCALED, bool FAST>
__global__ void __launch_bounds__(kSnipThreads, (KCH <= 4 ? 3 : (KCH <= 8 ? 2 : 1)))
snip_kernel(const T *__restrict__ spec, long long n_px, long long ld_spec, int ch0, int n_ch,
            const uint32_t *__restrict__ lohi, int ld_tab, int n_pass, int boxcar, float *__restrict__ out,
            long long ld_out, int out_mode, int *__restrict__ overflow, int prefetch_ahead) {
    extern __shared__ float4 sb[];  // [n_ch] x (4 pixels)
    const int tid = threadIdx.x;
    const long long px0 = (long long)blockIdx.x * kSnipPix;
    const int half = boxcar / 2;
    const float wgt = 1.0f / (float)boxcar;
    const T *rows[kSnipPix];
#pragma unroll
    for (int p = 0; p < kSnipPix; ++p) rows[p] = spec + (px0 + p < n_px ? px0 + p : px0) * ld_spec + ch0;

    float4 cur[KCH];

    // ---- stage raw counts in shared memory -------------------------------------------------
    // all global loads of the thread are issued before the first shared-memory store, so the CTA pays ONE
    // memory latency here, not one per channel (cur[] is free until the boxcar step)
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_ch) {
            r.x = load_count(rows[0] + i);
            r.y = (px0 + 1 < n_px) ? load_count(rows[1] + i) : 0.f;
            r.z = (px0 + 2 < n_px) ? load_count(rows[2] + i) : 0.f;
            r.w = (px0 + 3 < n_px) ? load_count(rows[3] + i) : 0.f;
        }
        cur[k] = r;
    }
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        if (i < n_ch) sb[i] = cur[k];
    }
    __syncthreads();

    // ---- boxcar mean (zero padded, bkg.py:93-97) and log1p (bkg.py:100-101) -----------------
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_ch) {
            if (boxcar == 5 && i >= 2 && i + 2 < n_ch) {
                // the default window, away from the edges: five unconditional 128-bit loads
#pragma unroll
                for (int j = -2; j <= 2; ++j) {
                    axpy4(acc, sb[i + j], wgt);
                }
            } else {
                for (int j = -half; j <= half; ++j) {
                    const int q = i + j;
                    if (q >= 0 && q < n_ch) {  // zero padding contributes +0
                        const float4 v = sb[q];
                        acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, wgt));
                        acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, wgt));
                        acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, wgt));
                        acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, wgt));
                    }
                }
            }
            acc.x = finite_or_zero(log_1p<FAST>(acc.x));
            acc.y = finite_or_zero(log_1p<FAST>(acc.y));
            acc.z = finite_or_zero(log_1p<FAST>(acc.z));
            acc.w = finite_or_zero(log_1p<FAST>(acc.w));
        }
        cur[k] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        if (i < n_ch) sb[i] = cur[k];
    }

    // ---- clipping passes (bkg.py:53-66, 104-112), Jacobi style ------------------------------
    uint32_t tab[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        if (SCALED)
            tab[k] = __ldg(lohi + i);
        else
            tab[k] = (n_pass > 0 && i < n_ch) ? __ldg(lohi + i) : 0u;
    }
    __syncthreads();
    if (prefetch_ahead > 0) {
        // pull the rows of the CTA that will follow this one on the SM into L2 while the passes run: its prologue then
        // waits for L2 instead of HBM (the load phase held 20 % of this kernel's stall samples, 69 % of them on memory)
        const long long qx = ((long long)blockIdx.x + prefetch_ahead) * kSnipPix;
        const long long bytes_per_row = (long long)n_ch * (long long)sizeof(T);
        const long long lines_per_row = (bytes_per_row + 127) / 128;
        for (long long l = tid; l < lines_per_row * kSnipPix; l += kSnipThreads) {
            const long long p = l / lines_per_row, off = (l - p * lines_per_row) * 128;
            if (qx + p < n_px) {
                const char *addr = reinterpret_cast<const char *>(spec + (qx + p) * ld_spec + ch0) + off;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
            }
        }
    }
    const uint32_t sb_base = static_cast<uint32_t>(__cvta_generic_to_shared(sb));
    for (int pass = 0; pass < n_pass; ++pass) {
        const uint32_t *next = lohi + (long long)(pass + 1) * ld_tab;
        const bool more = pass + 1 < n_pass;
        // channels past n_ch carry index word 0: they read channel 0 and their result is never stored
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            const int i = tid + k * kSnipThreads;
            const float4 a = lds128(sb_base + (SCALED ? (tab[k] & 0xffffu) : ((tab[k] & 0xffffu) << 4)));
            const float4 b = lds128(sb_base + (SCALED ? (tab[k] >> 16) : ((tab[k] >> 12) & 0xffff0u)));
            // this word is consumed: fetch the next pass's (from L2) behind the rest of the pass
            tab[k] = (more && i < n_ch) ? __ldg(next + i) : 0u;
            clip4(a, b, cur[k]);
        }
        if (!more) break;  // the result of the last pass stays in registers
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            const int i = tid + k * kSnipThreads;
            if (i < n_ch) sts128(sb_base + ((uint32_t)i << 4), cur[k]);
        }
        __syncthreads();
    }

    // ---- back to counts (bkg.py:113-114) and output -----------------------------------------
    // the counts of channel k + 1 are fetched while channel k goes through expf and the stores
    const bool need_y = out_mode != MT_SNIP_BACKGROUND;
    float y_next[kSnipPix];
    auto fetch = [&](int k) {
        const int i = tid + k * kSnipThreads;
#pragma unroll
        for (int p = 0; p < kSnipPix; ++p) y_next[p] = (need_y && i < n_ch) ? load_count(rows[p] + i) : 0.f;
    };
    fetch(0);
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        const int i = tid + k * kSnipThreads;
        float y[kSnipPix];
#pragma unroll
        for (int p = 0; p < kSnipPix; ++p) y[p] = y_next[p];
        if (k + 1 < KCH) fetch(k + 1);
        if (i >= n_ch) continue;
        const float vals[kSnipPix] = {cur[k].x, cur[k].y, cur[k].z, cur[k].w};
#pragma unroll
        for (int p = 0; p < kSnipPix; ++p) {
            const long long px = px0 + p;
            if (px >= n_px) continue;
            float *orow = out + px * ld_out;
            const float bkg = finite_or_zero(exp_m1<FAST>(vals[p]));
            if (out_mode == MT_SNIP_RESIDUAL_HILO_F16) {
                // two fp16 planes (4 bytes per channel instead of 8): hi = r rounded to fp16, lo = the remainder; the
                // contraction then runs on the fp16 tensor-core path.  |r| beyond the fp16 range raises the flag.
                __half *hrow = reinterpret_cast<__half *>(out) + px * ld_out;
                const float r = __fsub_rn(y[p], bkg);
                const __half hi = __float2half_rn(r);
                hrow[i] = hi;
                hrow[n_ch + i] = __float2half_rn(__fsub_rn(r, __half2float(hi)));
                if (!(fabsf(r) <= 65504.f)) atomicOr(overflow, 1);
            } else if (out_mode == MT_SNIP_BACKGROUND) {
                orow[i] = bkg;
            } else {
                const float r = __fsub_rn(y[p], bkg);
                if (out_mode == MT_SNIP_RESIDUAL) {
                    orow[i] = r;
                } else {
                    const float hi = mt::tf32_round(r);
                    orow[i] = hi;
                    orow[n_ch + i] = __fsub_rn(r, hi);
                }
            }
        }
    }
}

// tuning knobs (environment): MT_SNIP_PREFETCH = CTAs ahead whose rows are prefetched into L2 (default 0 = off;
// 2 x SM count = the CTA that takes this one's place), MT_SNIP_FASTMATH (default 1 since the contiguous-run kernel, where
// it is worth 10 %; 0 = the 1-ulp logf / expf) selects MUFU-based log / exp in BOTH kernels.  Measured on
// B200 (c3 shard, 131 072 px): neither moves the kernel (1.091 ms per 32 768 px without, 1.095 with prefetch, 1.109 with
// fast math): with two resident CTAs the phases that wait on memory or issue math are already hidden behind the other
// CTA's pass loop, and the pass loop itself runs at the shared-memory rate one CTA can sustain between barriers.
int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v ? atoi(v) : fallback;
}

template <typename T, int KCH, bool SCALED, bool FAST>
int launch_snip(const void *spec, int64_t n_px, int64_t ld_spec, int ch0, int n_ch, const uint32_t *lohi, int ld_tab,
                int n_pass, int boxcar, float *out, int64_t ld_out, int out_mode, int *overflow, cudaStream_t stream) {
    auto kern = snip_kernel<T, KCH, SCALED, FAST>;
    static const int prefetch = env_int("MT_SNIP_PREFETCH", 0);
    const size_t smem = sizeof(float4) * (size_t)n_ch;
    static int smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 16) smem_set[dev] = (int)smem;
    }
    const long long n_cta = (n_px + kSnipPix - 1) / kSnipPix;
    kern<<<(unsigned)n_cta, kSnipThreads, smem, stream>>>(static_cast<const T *>(spec), n_px, ld_spec, ch0, n_ch, lohi,
                                                          ld_tab, n_pass, boxcar, out, ld_out, out_mode, overflow, prefetch);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

// rows of a scaled table: n_ch rounded up to the thread grid of the kernel variant that handles it
int scaled_pitch(int n_ch) {
    for (int k = 2; k <= 8; k *= 2)
        if (n_ch <= kSnipThreads * k) return kSnipThreads * k;
    return 0;  // wider windows keep plain indices (their byte offsets do not fit 16 bits)
}

template <typename T>
int dispatch_snip(const void *spec, int64_t n_px, int64_t ld_spec, int ch0, int n_ch, const uint32_t *lohi, int ld_tab,
                  bool scaled, int n_pass, int boxcar, float *out, int64_t ld_out, int out_mode, int *overflow,
                  cudaStream_t stream) {
    static const bool fast = env_int("MT_SNIP_FASTMATH", 1) != 0;
#define MT_SNIP_ARGS spec, n_px, ld_spec, ch0, n_ch, lohi, ld_tab, n_pass, boxcar, out, ld_out, out_mode, overflow, stream
#define MT_SNIP_CASE(K)                                                                 \
    if (n_ch <= kSnipThreads * K) {                                                     \
        if (scaled && fast) return launch_snip<T, K, true, true>(MT_SNIP_ARGS);         \
        if (scaled) return launch_snip<T, K, true, false>(MT_SNIP_ARGS);                \
        if (fast) return launch_snip<T, K, false, true>(MT_SNIP_ARGS);                  \
        return launch_snip<T, K, false, false>(MT_SNIP_ARGS);                           \
    }
    MT_SNIP_CASE(2)
    MT_SNIP_CASE(4)
    MT_SNIP_CASE(8)
#undef MT_SNIP_CASE
    if (n_ch <= kSnipThreads * 16 && !scaled) {
        if (fast) return launch_snip<T, 16, false, true>(MT_SNIP_ARGS);
        return launch_snip<T, 16, false, false>(MT_SNIP_ARGS);
    }
#undef MT_SNIP_ARGS
    return mt::fail(MT_ERR_UNSUPPORTED, "mt_snip: n_ch=%d exceeds the on-chip limit of %d channels", n_ch,
                    kSnipThreads * (scaled ? 8 : 16));
}

__global__ void snip_scale_tables_kernel(const uint32_t *__restrict__ lohi, int n_pass, int n_ch, uint32_t *__restrict__ out,
                                         int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, pass = blockIdx.y;
    if (i >= ld) return;
    uint32_t w = 0u;
    if (pass < n_pass && i < n_ch) {
        const uint32_t v = lohi[(long long)pass * n_ch + i];
        w = ((v & 0xffffu) << 4) | ((v >> 16) << 20);
    }
    out[(long long)pass * ld + i] = w;
}

int snip_common(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0, int32_t n_ch,
                const uint32_t *tab_dev, int ld_tab, bool scaled, int32_t n_pass, int32_t boxcar, float *out_dev,
                int64_t ld_out, int32_t out_mode, int *overflow, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && tab_dev && out_dev, "mt_snip: null pointer");
    MT_REQUIRE(n_px >= 0 && n_ch > 0 && ch0 >= 0 && ld_spec >= (int64_t)ch0 + n_ch,
               "mt_snip: bad sizes n_px=%lld n_ch=%d ch0=%d ld=%lld", (long long)n_px, n_ch, ch0,
               (long long)ld_spec);
    MT_REQUIRE(n_pass >= 0, "mt_snip: n_pass=%d", n_pass);
    MT_REQUIRE(boxcar >= 1 && (boxcar & 1), "mt_snip: boxcar must be odd (got %d)", boxcar);
    MT_REQUIRE(out_mode >= MT_SNIP_BACKGROUND && out_mode <= MT_SNIP_RESIDUAL_HILO_F16, "mt_snip: out_mode=%d",
               out_mode);
    MT_REQUIRE(out_mode != MT_SNIP_RESIDUAL_HILO_F16 || overflow, "mt_snip: fp16 planes need an overflow flag");
    MT_REQUIRE(ld_out >= (out_mode >= MT_SNIP_RESIDUAL_HILO ? 2 : 1) * (int64_t)n_ch,
               "mt_snip: ld_out=%lld too small", (long long)ld_out);
    MT_REQUIRE(n_ch <= 65535, "mt_snip: n_ch=%d does not fit the 16-bit gather indices", n_ch);
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        // measurement knob: run only the first MT_SNIP_DEBUG_NPASS passes (results are then NOT the SNIP background);
        // T(14) - T(7) gives the marginal cost of a pass, T(0) the prologue + epilogue
        static const int cap = env_int("MT_SNIP_DEBUG_NPASS", -1);
        if (cap >= 0 && cap < n_pass) n_pass = cap;
    }
    if (dtype == MT_DTYPE_F32)
        return dispatch_snip<float>(spec_dev, n_px, ld_spec, ch0, n_ch, tab_dev, ld_tab, scaled, n_pass, boxcar, out_dev,
                                    ld_out, out_mode, overflow, stream);
    if (dtype == MT_DTYPE_F16)
        return dispatch_snip<__half>(spec_dev, n_px, ld_spec, ch0, n_ch, tab_dev, ld_tab, scaled, n_pass, boxcar, out_dev,
                                     ld_out, out_mode, overflow, stream);
    return mt::fail(MT_ERR_UNSUPPORTED, "mt_snip: dtype %d", dtype);
}

}  // namespace

extern "C" int mt_snip(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0,
                       int32_t n_ch, const uint32_t *lohi_dev, int32_t n_pass, int32_t boxcar,
                       float *out_dev, int64_t ld_out, int32_t out_mode, void *stream_) {
    MT_REQUIRE(out_mode != MT_SNIP_RESIDUAL_HILO_F16, "mt_snip: fp16 planes are written by mt_snip_residual_f16");
    return snip_common(spec_dev, dtype, n_px, ld_spec, ch0, n_ch, lohi_dev, n_ch, false, n_pass, boxcar, out_dev, ld_out,
                       out_mode, nullptr, stream_);
}

extern "C" int mt_snip_scaled_pitch(int32_t n_ch) { return scaled_pitch(n_ch); }

extern "C" int mt_snip_scale_tables(const uint32_t *lohi_dev, int32_t n_pass, int32_t n_ch, uint32_t *scaled_dev,
                                    int32_t ld_scaled, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(lohi_dev && scaled_dev, "mt_snip_scale_tables: null pointer");
    MT_REQUIRE(n_pass >= 0 && n_ch > 0, "mt_snip_scale_tables: bad sizes");
    MT_REQUIRE(scaled_pitch(n_ch) > 0 && ld_scaled == scaled_pitch(n_ch),
               "mt_snip_scale_tables: n_ch=%d needs a row pitch of %d words (0 = not available), got %d", n_ch,
               scaled_pitch(n_ch), ld_scaled);
    dim3 grid((ld_scaled + 255) / 256, n_pass + 1);  // one zero row after the last pass
    snip_scale_tables_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(lohi_dev, n_pass, n_ch, scaled_dev,
                                                                                   ld_scaled);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

// The gather table of every clipping pass built on the device from six fp32 calibration values: the operation sequence
// of bkg.py:83-91 (widths), 104-112 (pass schedule) and 57-62 (clamp, truncate) with explicitly rounded single-precision
// operations (no contraction into FMAs), i.e. bit-identical to the reference's torch CPU ops and to tables.snip_tables.
// fit_spec(tune_params=True) rebuilds this table at most of its trial points; on the host it cost 0.15 ms + a copy.
__global__ void snip_tables_f32_kernel(int n_ch, float ch0, float e_offset, float e_slope, float e_quad, float snip_width,
                                       float fwhm_offset, float fano, int n_pass, uint32_t *__restrict__ lohi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ch) return;
    const float idx = (float)i, top = (float)(n_ch - 1);
    const float g = __fadd_rn(idx, ch0);
    const float energy = __fadd_rn(__fadd_rn(e_offset, __fmul_rn(e_slope, g)), __fmul_rn(e_quad, __fmul_rn(g, g)));
    const float t = __fdiv_rn(fwhm_offset, 2.3548f);
    const float sigma_sq = fmaxf(__fadd_rn(__fmul_rn(t, t), __fmul_rn(__fmul_rn(energy, 2.96f), fano)), 0.f);
    float w = __fdiv_rn(__fmul_rn(__fmul_rn(snip_width, 2.35f), __fsqrt_rn(sigma_sq)), e_slope);
    for (int pass = 0; pass < n_pass; ++pass) {
        if (pass >= 3) w = __fdiv_rn(w, 1.41421356237309504880f);  // passes 0-2 use the full width, then / sqrt(2) each
        const int lo = __float2int_rz(fminf(fmaxf(__fsub_rn(idx, w), 0.f), top));
        const int hi = __float2int_rz(fminf(fmaxf(__fadd_rn(idx, w), 0.f), top));
        lohi[(long long)pass * n_ch + i] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
}

extern "C" int mt_snip_tables_f32(int32_t n_ch, int32_t ch0, float e_offset, float e_slope, float e_quad, float snip_width,
                                  float fwhm_offset, float fwhm_fanoprime, int32_t n_pass, uint32_t *lohi_dev,
                                  void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(lohi_dev, "mt_snip_tables_f32: null pointer");
    MT_REQUIRE(n_ch > 0 && n_ch <= 65535 && n_pass >= 0 && n_pass <= 4096, "mt_snip_tables_f32: bad sizes");
    if (n_pass == 0) return MT_OK;
    snip_tables_f32_kernel<<<(n_ch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        n_ch, (float)ch0, e_offset, e_slope, e_quad, snip_width, fwhm_offset, fwhm_fanoprime, n_pass, lohi_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_snip_scaled(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0,
                              int32_t n_ch, const uint32_t *scaled_dev, int32_t ld_scaled, int32_t n_pass, int32_t boxcar,
                              float *out_dev, int64_t ld_out, int32_t out_mode, void *stream_) {
    MT_REQUIRE(scaled_pitch(n_ch) > 0 && ld_scaled == scaled_pitch(n_ch), "mt_snip_scaled: table pitch %d for n_ch=%d",
               ld_scaled, n_ch);
    MT_REQUIRE(out_mode != MT_SNIP_RESIDUAL_HILO_F16, "mt_snip_scaled: fp16 planes are written by mt_snip_residual_f16");
    return snip_common(spec_dev, dtype, n_px, ld_spec, ch0, n_ch, scaled_dev, ld_scaled, true, n_pass, boxcar, out_dev,
                       ld_out, out_mode, nullptr, stream_);
}

extern "C" int mt_snip_residual_f16(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0,
                                    int32_t n_ch, const uint32_t *tab_dev, int32_t ld_tab, int32_t scaled, int32_t n_pass,
                                    int32_t boxcar, void *out_half_dev, int64_t ld_out, int32_t *overflow_dev, void *stream_) {
    if (scaled)
        MT_REQUIRE(scaled_pitch(n_ch) > 0 && ld_tab == scaled_pitch(n_ch), "mt_snip_residual_f16: table pitch %d for n_ch=%d",
                   ld_tab, n_ch);
    return snip_common(spec_dev, dtype, n_px, ld_spec, ch0, n_ch, tab_dev, scaled ? ld_tab : n_ch, scaled != 0, n_pass, boxcar,
                       static_cast<float *>(out_half_dev), ld_out, MT_SNIP_RESIDUAL_HILO_F16, overflow_dev, stream_);
}
