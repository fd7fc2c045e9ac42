// K5 -- amplitude-sparsity penalty of the MSE objective (opt.py:372-396).
//
// The reference adds  l1_lambda * loss(bkg, y) / n_elements * sum_e A_e  to the loss.  With A >= 0 the
// term is linear, so the penalised fit is the non-negative least-squares problem with right-hand side
// h - c/2, i.e. the unconstrained amplitudes move by  -(c_p / 2) * (G^-1 1)_e  where
// c_p = l1_lambda / n_elements * sum_c (y[p,c] - bkg[p,c])^2  is the only per-pixel quantity.  This kernel
// computes that row reduction from the operand the contraction just consumed (one more streaming read of
// the spectra -- the price of the option) and applies the shift to x in place; the sign-constrained solve
// (mt_nnls_fixup) follows.  One warp per pixel, 16-byte loads when the rows allow it.
#include <cuda_fp16.h>

#include "mt_common.cuh"

namespace {

__device__ __forceinline__ float as_float(float v) { return v; }
__device__ __forceinline__ float as_float(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
penalty_shift_kernel(const T *__restrict__ spec, long long n_px, long long ld_spec, int col0, int n_ch, int repeat,
                     const float *__restrict__ chan_mask, const float *__restrict__ coef, float *__restrict__ x,
                     long long ld_x, int pixel_major, int n_comp, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    constexpr int kVec = 16 / (int)sizeof(T);
    for (long long p = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < n_px; p += warps) {
        const T *row = spec + p * ld_spec + col0;
        float acc = 0.f;
        if (repeat == 1 && chan_mask == nullptr && vec_ok) {
            // plain streamed rows: 128-bit loads
            const uint4 *v = reinterpret_cast<const uint4 *>(row);
            for (int i = lane; i < n_ch / kVec; i += 32) {
                const uint4 q = __ldg(v + i);
                const T *e = reinterpret_cast<const T *>(&q);
#pragma unroll
                for (int k = 0; k < kVec; ++k) {
                    const float f = as_float(e[k]);
                    acc = fmaf(f, f, acc);
                }
            }
        } else {
            for (int c = lane; c < n_ch; c += 32) {
                float f = as_float(__ldg(row + c));
                for (int r = 1; r < repeat; ++r) f += as_float(__ldg(row + (long long)r * n_ch + c));  // hi + remainder
                const float m = chan_mask ? chan_mask[c] : 1.f;
                acc = fmaf(m * f, f, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        for (int e = lane; e < n_comp; e += 32) {
            float *dst = pixel_major ? x + p * ld_x + e : x + (long long)e * ld_x + p;
            *dst = fmaf(-acc, coef[e], *dst);
        }
    }
}

}  // namespace

extern "C" int mt_penalty_shift(const void *spec_dev, int32_t spec_dtype, int64_t n_px, int64_t ld_spec, int32_t col0,
                                int32_t n_ch, int32_t repeat, const float *chan_mask_dev, const float *coef_dev,
                                float *x_dev, int64_t ld_x, int32_t x_layout, int32_t n_comp, void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && coef_dev && x_dev, "mt_penalty_shift: null pointer");
    MT_REQUIRE(n_px >= 0 && n_ch > 0 && repeat >= 1 && col0 >= 0 && n_comp > 0, "mt_penalty_shift: bad sizes");
    MT_REQUIRE(spec_dtype == MT_DTYPE_F32 || spec_dtype == MT_DTYPE_F16, "mt_penalty_shift: dtype %d", spec_dtype);
    MT_REQUIRE(x_layout == MT_X_ELEMENT_MAJOR || x_layout == MT_X_PIXEL_MAJOR, "mt_penalty_shift: layout %d", x_layout);
    if (n_px == 0) return MT_OK;
    const size_t es = spec_dtype == MT_DTYPE_F16 ? 2 : 4;
    const int vec_ok = ((uintptr_t)spec_dev % 16 == 0) && ((ld_spec * es) % 16 == 0) && ((col0 * es) % 16 == 0) &&
                       ((n_ch * es) % 16 == 0);
    const long long blocks_needed = (n_px + 7) / 8;  // 8 warps per block
    const long long cap = (long long)mt::num_sms() * 16;
    const unsigned grid = (unsigned)(blocks_needed < cap ? blocks_needed : cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int pm = x_layout == MT_X_PIXEL_MAJOR;
    if (spec_dtype == MT_DTYPE_F16)
        penalty_shift_kernel<__half><<<grid, 256, 0, s>>>(static_cast<const __half *>(spec_dev), n_px, ld_spec, col0,
                                                          n_ch, repeat, chan_mask_dev, coef_dev, x_dev, ld_x, pm,
                                                          n_comp, vec_ok);
    else
        penalty_shift_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(spec_dev), n_px, ld_spec, col0,
                                                         n_ch, repeat, chan_mask_dev, coef_dev, x_dev, ld_x, pm,
                                                         n_comp, vec_ok);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

// ---- exact operand split for fp32 volumes whose counts exceed 11 significant bits ------------------------------
// out[p][0 .. n_ch)      = y rounded to nearest to 11 significant bits (exactly representable as a tf32 operand)
// out[p][n_ch .. 2 n_ch) = y - hi (exact in fp32; its own top 11 bits carry the result to 22 bits)
// One read of the window, two writes: the layout K3 contracts with `repeat = 2`, as the SNIP kernel emits it.
namespace {

__global__ void __launch_bounds__(256)
split_hilo_kernel(const float *__restrict__ spec, long long n_px, long long ld_spec, int col0, int n_ch,
                  float *__restrict__ out, long long ld_out) {
    const long long total = n_px * (long long)n_ch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / n_ch;
        const int c = (int)(i - p * n_ch);
        const float y = __ldg(spec + p * ld_spec + col0 + c);
        const float hi = mt::tf32_round(y);
        out[p * ld_out + c] = hi;
        out[p * ld_out + n_ch + c] = __fsub_rn(y, hi);
    }
}

// flag[0] |= 1 when any value of the window is not exactly representable as a tf32 operand (low 13 mantissa bits set;
// NaN / Inf included): integer counts >= 2048, but also normalised or dead-time corrected data.  One read of the window.
__global__ void __launch_bounds__(256)
operand_inexact_kernel(const float *__restrict__ spec, long long n_px, long long ld_spec, int col0, int n_ch,
                       int *__restrict__ flag) {
    const long long total = n_px * (long long)n_ch;
    unsigned bad = 0u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / n_ch;
        const int c = (int)(i - p * n_ch);
        const unsigned b = __float_as_uint(__ldg(spec + p * ld_spec + col0 + c));
        bad |= (b & 0x1fffu) | (((b >> 23) & 0xffu) == 0xffu ? 1u : 0u);
    }
    if (__any_sync(0xffffffffu, bad != 0u) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

}  // namespace

extern "C" int mt_operand_inexact(const float *spec_dev, int64_t n_px, int64_t ld_spec, int32_t col0, int32_t n_ch,
                                  int32_t *flag_dev, void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && flag_dev, "mt_operand_inexact: null pointer");
    MT_REQUIRE(n_px >= 0 && n_ch > 0 && col0 >= 0 && ld_spec >= (int64_t)col0 + n_ch, "mt_operand_inexact: bad sizes");
    if (n_px == 0) return MT_OK;
    const long long total = n_px * (long long)n_ch;
    const long long want = (total + 255) / 256, cap = (long long)mt::num_sms() * 32;
    operand_inexact_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        spec_dev, n_px, ld_spec, col0, n_ch, flag_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_split_hilo(const float *spec_dev, int64_t n_px, int64_t ld_spec, int32_t col0, int32_t n_ch,
                             float *out_dev, int64_t ld_out, void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && out_dev, "mt_split_hilo: null pointer");
    MT_REQUIRE(n_px >= 0 && n_ch > 0 && col0 >= 0 && ld_spec >= (int64_t)col0 + n_ch && ld_out >= 2 * (int64_t)n_ch,
               "mt_split_hilo: bad sizes");
    if (n_px == 0) return MT_OK;
    const long long total = n_px * (long long)n_ch;
    const long long want = (total + 255) / 256, cap = (long long)mt::num_sms() * 32;
    split_hilo_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        spec_dev, n_px, ld_spec, col0, n_ch, out_dev, ld_out);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
