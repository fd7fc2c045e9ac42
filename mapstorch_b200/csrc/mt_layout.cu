// Data-format steps either side of the fit: channel-first volumes and the integrated spectrum.
//
//  * mt_transpose_cf : raw XRF-Maps files store MAPS/Spectra/mca_arr channel-first, [C][H][W]; the reference's
//                      create_dataset moves the energy axis last on the host with np.moveaxis (io.py:154-156).
//                      Here a row slab [C][rows*W] that has just crossed PCIe is transposed to the [P][C]
//                      layout the fit streams, on the device: 32 x 32 shared-memory tiles, coalesced on both sides.
//  * mt_int_spec     : the integrated spectrum int_spec[c] = sum over pixels (io.py:159), accumulated slab by slab as
//                      a by-product of the streaming pass.  fp64 accumulators: sums of integer counts are exact, so
//                      the result does not depend on the order of the atomics (deterministic, equal to the host sum).
// Both are HBM-bound single-pass kernels.
#include <cuda_fp16.h>

#include "mt_common.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
transpose_cf_kernel(const T *__restrict__ src, long long ld_src, int n_ch, long long n_px, T *__restrict__ dst,
                    long long ld_dst) {
    __shared__ T tile[32][33];
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int c = c0 + ty + j;
        const long long p = p0 + tx;
        if (c < n_ch && p < n_px) tile[ty + j][tx] = src[(long long)c * ld_src + p];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const long long p = p0 + ty + j;
        const int c = c0 + tx;
        if (c < n_ch && p < n_px) dst[p * ld_dst + c] = tile[tx][ty + j];
    }
}

template <typename T>
__device__ __forceinline__ double as_double(T v);
template <>
__device__ __forceinline__ double as_double<float>(float v) { return (double)v; }
template <>
__device__ __forceinline__ double as_double<__half>(__half v) { return (double)__half2float(v); }

constexpr int kIntSpecRows = 256;  // pixels per CTA

template <typename T>
__global__ void __launch_bounds__(128)
int_spec_kernel(const T *__restrict__ spec, long long n_px, long long ld, int n_ch, double *__restrict__ acc) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    const long long p_begin = (long long)blockIdx.y * kIntSpecRows;
    const long long p_end = p_begin + kIntSpecRows < n_px ? p_begin + kIntSpecRows : n_px;
    if (c >= n_ch) return;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four independent chains: loads in flight
    long long p = p_begin;
    for (; p + 3 < p_end; p += 4) {
        s0 += as_double<T>(__ldg(spec + p * ld + c));
        s1 += as_double<T>(__ldg(spec + (p + 1) * ld + c));
        s2 += as_double<T>(__ldg(spec + (p + 2) * ld + c));
        s3 += as_double<T>(__ldg(spec + (p + 3) * ld + c));
    }
    for (; p < p_end; ++p) s0 += as_double<T>(__ldg(spec + p * ld + c));
    atomicAdd(acc + c, (s0 + s1) + (s2 + s3));
}

}  // namespace

extern "C" int mt_transpose_cf(const void *src_dev, int32_t dtype, int32_t n_ch, int64_t n_px, int64_t ld_src, void *dst_dev,
                               int64_t ld_dst, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(src_dev && dst_dev, "mt_transpose_cf: null pointer");
    MT_REQUIRE(n_ch > 0 && n_px >= 0 && ld_src >= n_px && ld_dst >= n_ch, "mt_transpose_cf: sizes n_ch=%d n_px=%lld ld_src=%lld "
               "ld_dst=%lld", n_ch, (long long)n_px, (long long)ld_src, (long long)ld_dst);
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long bx = (n_px + 31) / 32;
    MT_REQUIRE(bx < (1ll << 31), "mt_transpose_cf: n_px=%lld", (long long)n_px);
    dim3 grid((unsigned)bx, (unsigned)((n_ch + 31) / 32));
    if (dtype == MT_DTYPE_F32)
        transpose_cf_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(src_dev), ld_src, n_ch, n_px,
                                                             static_cast<float *>(dst_dev), ld_dst);
    else if (dtype == MT_DTYPE_F16)
        transpose_cf_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(src_dev), ld_src, n_ch, n_px,
                                                              static_cast<__half *>(dst_dev), ld_dst);
    else
        return mt::fail(MT_ERR_UNSUPPORTED, "mt_transpose_cf: dtype %d", dtype);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

extern "C" int mt_int_spec(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t n_ch, double *acc_dev,
                           void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && acc_dev, "mt_int_spec: null pointer");
    MT_REQUIRE(n_ch > 0 && n_px >= 0 && ld_spec >= n_ch, "mt_int_spec: sizes");
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long by = (n_px + kIntSpecRows - 1) / kIntSpecRows;
    MT_REQUIRE(by <= 65535ll * 64, "mt_int_spec: n_px=%lld (split the call)", (long long)n_px);
    // gridDim.y is limited to 65535: larger slabs are summed in several launches
    for (long long y0 = 0; y0 < by; y0 += 65535) {
        const long long ny = by - y0 < 65535 ? by - y0 : 65535;
        const long long p0 = y0 * kIntSpecRows;
        const long long np = n_px - p0 < ny * kIntSpecRows ? n_px - p0 : ny * kIntSpecRows;
        dim3 grid((unsigned)((n_ch + 127) / 128), (unsigned)ny);
        if (dtype == MT_DTYPE_F32)
            int_spec_kernel<float><<<grid, 128, 0, stream>>>(static_cast<const float *>(spec_dev) + p0 * ld_spec, np, ld_spec,
                                                              n_ch, acc_dev);
        else if (dtype == MT_DTYPE_F16)
            int_spec_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half *>(spec_dev) + p0 * ld_spec, np, ld_spec,
                                                               n_ch, acc_dev);
        else
            return mt::fail(MT_ERR_UNSUPPORTED, "mt_int_spec: dtype %d", dtype);
    }
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
