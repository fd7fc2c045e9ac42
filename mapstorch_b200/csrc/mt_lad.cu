// Per-pixel loss="l1" maps (opt.py:232-233, 371-375: sum of absolute residuals instead of squares).
//
// With the calibration fixed the problem per pixel is  min_{x >= 0} sum_c |sum_e B[e,c] x_e - t_c|,  t = y - bkg: a
// linear programme.  It is solved for all pixels at once by ADMM on the splitting  z = B^T x - t:
//     x  <- argmin_{x >= 0} |B^T x - (t + z - u)|^2      the ordinary map fit of a modified target: the SAME streaming
//                                                       tensor-core contraction with the pixel-independent pseudo-inverse
//                                                       (mt_fit_contract) and non-negative fix-up as the MSE fit
//     z  <- soft(B^T x - t + u, kappa_p)                 |
//     u  <- u + B^T x - t - z                            |  this kernel, fused with the model evaluation B^T x and
//     q  <- t + z - u;  dq = q - q_previous  (hi | lo)   |  the operand split the contraction wants
// The x-update is linear up to the non-negative fix-up, so the contraction is applied to the CHANGE dq of the target and
// the unconstrained solution is accumulated (xu += W dq): the rounding of the tensor-core pass (22-bit operands) is then
// relative to dq, which vanishes as the iteration converges, instead of to q (an error floor of ~3e-5 in the objective).
// kappa_p (the threshold 1 / rho) is per pixel: the x-update does not depend on it.  The kernel also returns the L1
// objective of the x it was given, so the host loop keeps the best iterate per pixel.
// One CTA = 8 warps = 16 pixels; channels in chunks of 128 with the basis chunk staged in shared memory (read once per
// CTA and chunk, used by 16 pixels); a lane owns channels c0 + lane + 32 j.  fp32 CUDA-core arithmetic: the model needs
// full precision and only E FMAs per (pixel, channel).
#include "mt_common.cuh"

namespace {

constexpr int kLadThreads = 256;
constexpr int kLadChunk = 128;
constexpr int kLadPixPerWarp = 2;
constexpr int kLadPix = (kLadThreads / 32) * kLadPixPerWarp;

__global__ void __launch_bounds__(kLadThreads)
lad_update_kernel(const float *__restrict__ t, long long ld_t, const float *__restrict__ x, long long ld_x,
                  const float *__restrict__ basis, long long ld_b, int E, int n_ch, long long n_px, float *__restrict__ u,
                  long long ld_u, const float *__restrict__ kappa, float *__restrict__ qstate, long long ld_qs,
                  float *__restrict__ q, long long ld_q, float *__restrict__ obj, int mode) {
    extern __shared__ float smem[];
    float *bs = smem;                      // [E][kLadChunk]
    float *xs = smem + (size_t)E * kLadChunk;  // [kLadPix][E]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long px0 = (long long)blockIdx.x * kLadPix;
    for (int i = tid; i < kLadPix * E; i += kLadThreads) {
        const int p = i / E, e = i - p * E;
        xs[i] = px0 + p < n_px ? x[(long long)e * ld_x + px0 + p] : 0.f;
    }
    const long long pa = px0 + warp * kLadPixPerWarp, pb = pa + 1;
    const bool va = pa < n_px, vb = pb < n_px;
    const float ka = (mode && va) ? kappa[pa] : 0.f, kb = (mode && vb) ? kappa[pb] : 0.f;
    const float *xa = xs + (warp * kLadPixPerWarp) * E, *xb = xa + E;
    float oa = 0.f, ob = 0.f;
    for (int c0 = 0; c0 < n_ch; c0 += kLadChunk) {
        __syncthreads();
        for (int i = tid; i < E * kLadChunk; i += kLadThreads) {
            const int e = i / kLadChunk, c = c0 + (i - e * kLadChunk);
            bs[i] = c < n_ch ? __ldg(basis + (long long)e * ld_b + c) : 0.f;
        }
        __syncthreads();
        float ma[4] = {0.f, 0.f, 0.f, 0.f}, mb[4] = {0.f, 0.f, 0.f, 0.f};
        for (int e = 0; e < E; ++e) {
            const float va_e = xa[e], vb_e = xb[e];
            const float *row = bs + e * kLadChunk + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float b = row[32 * j];
                ma[j] = fmaf(va_e, b, ma[j]);
                mb[j] = fmaf(vb_e, b, mb[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + lane + 32 * j;
            if (c >= n_ch) continue;
            if (va) {
                const float r = ma[j] - t[pa * ld_t + c];
                oa += fabsf(r);
                if (mode) {
                    const float v = r + u[pa * ld_u + c];
                    const float z = copysignf(fmaxf(fabsf(v) - ka, 0.f), v);
                    const float un = v - z;  // u + r - z
                    u[pa * ld_u + c] = un;
                    const float qv = ma[j] - r + z - un;  // t + z - u
                    const float dq = qv - qstate[pa * ld_qs + c];  // the x-update works on the CHANGE of the target
                    qstate[pa * ld_qs + c] = qv;
                    const float hi = mt::tf32_round(dq);
                    q[pa * ld_q + c] = hi;
                    q[pa * ld_q + n_ch + c] = __fsub_rn(dq, hi);
                }
            }
            if (vb) {
                const float r = mb[j] - t[pb * ld_t + c];
                ob += fabsf(r);
                if (mode) {
                    const float v = r + u[pb * ld_u + c];
                    const float z = copysignf(fmaxf(fabsf(v) - kb, 0.f), v);
                    const float un = v - z;
                    u[pb * ld_u + c] = un;
                    const float qv = mb[j] - r + z - un;
                    const float dq = qv - qstate[pb * ld_qs + c];
                    qstate[pb * ld_qs + c] = qv;
                    const float hi = mt::tf32_round(dq);
                    q[pb * ld_q + c] = hi;
                    q[pb * ld_q + n_ch + c] = __fsub_rn(dq, hi);
                }
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        oa += __shfl_xor_sync(0xffffffffu, oa, s);
        ob += __shfl_xor_sync(0xffffffffu, ob, s);
    }
    if (lane == 0 && obj) {
        if (va) obj[pa] = oa;
        if (vb) obj[pb] = ob;
    }
}

}  // namespace

extern "C" int mt_lad_update(const float *t_dev, int64_t ld_t, const float *x_dev, int64_t ld_x, const float *basis_dev,
                             int64_t ld_b, int32_t n_comp, int32_t n_ch, int64_t n_px, float *u_dev, int64_t ld_u,
                             const float *kappa_dev, float *qstate_dev, int64_t ld_qs, float *q_dev, int64_t ld_q,
                             float *obj_dev, int32_t mode, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(t_dev && x_dev && basis_dev, "mt_lad_update: null pointer");
    MT_REQUIRE(n_comp >= 1 && n_comp <= 104 && n_ch > 0 && n_px >= 0 && ld_t >= n_ch && ld_b >= n_ch && ld_x >= n_px,
               "mt_lad_update: sizes");
    MT_REQUIRE(mode == 0 || (u_dev && kappa_dev && q_dev && qstate_dev && ld_u >= n_ch && ld_qs >= n_ch &&
                             ld_q >= 2 * (int64_t)n_ch),
               "mt_lad_update: the update needs u, kappa, the target state and the operand planes (ld_q >= 2 n_ch)");
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t smem = sizeof(float) * ((size_t)n_comp * kLadChunk + (size_t)kLadPix * n_comp);
    static int smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(lad_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 16) smem_set[dev] = (int)smem;
    }
    const long long grid = (n_px + kLadPix - 1) / kLadPix;
    MT_REQUIRE(grid < (1ll << 31), "mt_lad_update: n_px=%lld", (long long)n_px);
    lad_update_kernel<<<(unsigned)grid, kLadThreads, smem, stream>>>(t_dev, ld_t, x_dev, ld_x, basis_dev, ld_b, n_comp, n_ch, n_px,
                                                                     u_dev, ld_u, kappa_dev, qstate_dev, ld_qs, q_dev, ld_q, obj_dev,
                                                                     mode);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
