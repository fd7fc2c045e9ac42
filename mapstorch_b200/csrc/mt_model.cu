// K1 -- fused MAPS forward-model kernel.
//
// Evaluates every additive term of the model (K/L/M Gaussians with step and tail terms,
// elastic and Compton peaks) per energy channel into one column per fitted component,
// replacing the ~10 ATen calls per line of the reference (map.py:57-124, 152-257).
// Arithmetic is fp32 in the reference's operation order with contraction disabled
// (__fmul_rn/__fadd_rn/__fdiv_rn), so columns agree with torch-CPU to a few ulp of
// expf/erfcf.  Bound: MUFU (one ex2 per Gaussian term and channel, erfc adds ~20 FP32 ops).
#include "mt_common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float eval_term(const MtTerm &t, float ev) {
    float d = __fsub_rn(ev, t.e);
    if (t.sign < 0.f) d = -d;
    if (t.kind == MT_TERM_GAUSS) {
        float q = __fdiv_rn(d, t.p0);
        float g = expf(__fmul_rn(-0.5f, __fmul_rn(q, q)));
        return __fmul_rn(t.amp, __fmul_rn(t.p1, g));
    }
    if (t.kind == MT_TERM_STEP) {
        float s = erfcf(__fdiv_rn(d, t.p0));
        return __fmul_rn(t.amp, __fmul_rn(t.p1, s));
    }
    // MT_TERM_TAIL
    float v = erfcf(__fadd_rn(__fdiv_rn(d, t.p0), t.p1));
    if (d < 0.f) v = __fmul_rn(expf(__fdiv_rn(d, t.p2)), v);
    return __fmul_rn(t.amp, __fmul_rn(t.p3, v));
}

// grid = (ceil(n_ch / kThreads), n_cols).  The terms of a column are staged through shared
// memory in slabs so that every thread reads them as broadcasts.
__global__ void __launch_bounds__(kThreads)
model_terms_kernel(const float *__restrict__ ev, int n_ch, const MtTerm *__restrict__ terms,
                   const int *__restrict__ col_start, float *__restrict__ out, long long ld_out) {
    __shared__ MtTerm slab[64];
    const int col = blockIdx.y;
    const int c = blockIdx.x * kThreads + threadIdx.x;
    const int t0 = col_start[col], t1 = col_start[col + 1];
    const float e = (c < n_ch) ? ev[c] : 0.f;
    float acc = 0.f;
    for (int base = t0; base < t1; base += 64) {
        const int n = min(64, t1 - base);
        __syncthreads();
        // 8 words per term: copy as a flat word array
        const uint32_t *src = reinterpret_cast<const uint32_t *>(terms + base);
        uint32_t *dst = reinterpret_cast<uint32_t *>(slab);
        for (int w = threadIdx.x; w < n * 8; w += kThreads) dst[w] = src[w];
        __syncthreads();
        for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, eval_term(slab[i], e));
    }
    if (c < n_ch) out[(long long)col * ld_out + c] = acc;
}

__global__ void __launch_bounds__(kThreads)
model_sum_kernel(const float *__restrict__ cols, long long ld, int n_cols, int n_ch, int bins,
                 float factor, float *__restrict__ out) {
    const int c = blockIdx.x * kThreads + threadIdx.x;
    if (c >= n_ch) return;
    float s = 0.f;
    for (int k = 0; k < n_cols; ++k) s = __fadd_rn(s, cols[(long long)k * ld + c]);
    if (bins > 0 && factor > 0.f) {
        float esc = 0.f;
        if (c + bins < n_ch) {
            float s2 = 0.f;
            for (int k = 0; k < n_cols; ++k) s2 = __fadd_rn(s2, cols[(long long)k * ld + c + bins]);
            esc = __fmul_rn(s2, factor);
        }
        s = __fadd_rn(s, esc);
    }
    out[c] = s;
}

// ev[c] = offset + slope * ch + quad * ch^2 for ch = ch0 + c, with the reference's rounding (map.py:271-281 evaluates
// it with separate fp32 torch ops: no fused multiply-add anywhere)
__global__ void energy_axis_kernel(int ch0, int n_ch, float off, float slope, float quad, float *__restrict__ ev) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_ch) return;
    const float ch = (float)(ch0 + c);
    ev[c] = __fadd_rn(__fadd_rn(off, __fmul_rn(slope, ch)), __fmul_rn(quad, __fmul_rn(ch, ch)));
}

}  // namespace

extern "C" int mt_energy_axis(int32_t ch0, int32_t n_ch, float e_offset, float e_slope, float e_quad, float *ev_dev,
                              void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(ev_dev && n_ch > 0 && ch0 >= 0 && (long long)ch0 + n_ch <= (1 << 24), "mt_energy_axis: bad arguments");
    energy_axis_kernel<<<(n_ch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(ch0, n_ch, e_offset, e_slope,
                                                                                            e_quad, ev_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_model_terms(const float *ev_dev, int32_t n_ch, const MtTerm *terms_host,
                              const int32_t *col_start_host, int32_t n_cols, float *out_dev,
                              int64_t ld_out, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(ev_dev && terms_host && col_start_host && out_dev, "mt_model_terms: null pointer");
    MT_REQUIRE(n_ch > 0 && n_cols > 0 && ld_out >= n_ch, "mt_model_terms: bad sizes n_ch=%d n_cols=%d ld=%lld",
               n_ch, n_cols, (long long)ld_out);
    MT_REQUIRE(col_start_host[0] == 0, "mt_model_terms: col_start[0] must be 0");
    for (int i = 0; i < n_cols; ++i)
        MT_REQUIRE(col_start_host[i + 1] >= col_start_host[i], "mt_model_terms: col_start not monotone");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int n_terms = col_start_host[n_cols];
    const size_t term_bytes = sizeof(MtTerm) * (size_t)(n_terms > 0 ? n_terms : 1);
    const size_t idx_bytes = sizeof(int32_t) * (size_t)(n_cols + 1);
    mt::StreamScratch holder;
    MT_CUDA_TRY(holder.alloc(term_bytes + idx_bytes, stream));
    char *scratch = holder.ptr;
    if (n_terms > 0)
        MT_CUDA_TRY(cudaMemcpyAsync(scratch, terms_host, sizeof(MtTerm) * (size_t)n_terms,
                                    cudaMemcpyHostToDevice, stream));
    MT_CUDA_TRY(cudaMemcpyAsync(scratch + term_bytes, col_start_host, idx_bytes, cudaMemcpyHostToDevice,
                                stream));
    dim3 grid((n_ch + kThreads - 1) / kThreads, n_cols);
    model_terms_kernel<<<grid, kThreads, 0, stream>>>(ev_dev, n_ch, reinterpret_cast<const MtTerm *>(scratch),
                                                      reinterpret_cast<const int *>(scratch + term_bytes),
                                                      out_dev, ld_out);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;  // the scratch goes back to the pool in stream order (StreamScratch)
}

extern "C" int mt_model_sum(const float *cols_dev, int64_t ld_cols, int32_t n_cols, int32_t n_ch,
                            int32_t escape_bins, float escape_factor, float *out_dev, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(cols_dev && out_dev, "mt_model_sum: null pointer");
    MT_REQUIRE(n_ch > 0 && n_cols >= 0 && ld_cols >= n_ch, "mt_model_sum: bad sizes");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    model_sum_kernel<<<(n_ch + kThreads - 1) / kThreads, kThreads, 0, stream>>>(
        cols_dev, ld_cols, n_cols, n_ch, escape_bins, escape_factor, out_dev);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
