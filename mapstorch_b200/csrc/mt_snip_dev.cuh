// Device helpers shared by the two SNIP kernels (mt_snip.cu: strided channel ownership, any shape;
// mt_snip_runs.cu: contiguous channel runs, persistent, for aligned volumes).  Both kernels perform the same fp32
// operations in the same order, so their results are bit-identical.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

#include "mt_common.cuh"
#include "mt_f32x2.cuh"

namespace mt_snip_dev {

__device__ __forceinline__ float load_count(const float *p) { return __ldg(p); }
__device__ __forceinline__ float load_count(const __half *p) { return __half2float(__ldg(p)); }

__device__ __forceinline__ float finite_or_zero(float v) {
    // torch.nan_to_num(nan=0, posinf=0, neginf=0)
    return (fabsf(v) <= 3.402823466e+38f) ? v : 0.f;
}

// log(1 + x) and exp(v) - 1 through the 1-ulp logf / expf.  The library's log1pf / expm1f keep RELATIVE accuracy
// for |x| << 1, which costs ~40 instructions each (30 % of this kernel's instruction count in the ncu source
// view); the parity bar of the background is relative to the row maximum, where the plain forms are as good:
// the rounding of 1 + x and of exp(v) - 1 is an absolute error of 6e-8 (x (1 + bkg)), far below 1e-5 of any row
// that holds at least one count.  Zero stays exactly zero.
template <bool FAST>
__device__ __forceinline__ float log_1p(float x) {
    // FAST: MUFU.LG2 * ln 2 (absolute error ~2e-7 for arguments in [1, 2^24]) instead of the 1-ulp logf
    return FAST ? __logf(__fadd_rn(1.f, x)) : logf(__fadd_rn(1.f, x));
}
template <bool FAST>
__device__ __forceinline__ float exp_m1(float v) {
    return __fsub_rn(FAST ? __expf(v) : expf(v), 1.f);
}

using mt::add2;
using mt::mul2;
using mt::pack2;
using mt::unpack2;

// cur <- min((a + b) / 2, cur) for the four pixels of a record
__device__ __forceinline__ void clip4(const float4 &a, const float4 &b, float4 &cur) {
    const unsigned long long half = pack2(0.5f, 0.5f);
    float m0, m1, m2, m3;
    unpack2(mul2(add2(pack2(a.x, a.y), pack2(b.x, b.y)), half), m0, m1);
    unpack2(mul2(add2(pack2(a.z, a.w), pack2(b.z, b.w)), half), m2, m3);
    cur.x = fminf(m0, cur.x);
    cur.y = fminf(m1, cur.y);
    cur.z = fminf(m2, cur.z);
    cur.w = fminf(m3, cur.w);
}

// acc <- acc + v * w for the four pixels of a record (two roundings per element, like the reference's conv1d)
__device__ __forceinline__ void axpy4(float4 &acc, const float4 &v, float w) {
    const unsigned long long w2 = pack2(w, w);
    unpack2(add2(pack2(acc.x, acc.y), mul2(pack2(v.x, v.y), w2)), acc.x, acc.y);
    unpack2(add2(pack2(acc.z, acc.w), mul2(pack2(v.z, v.w), w2)), acc.z, acc.w);
}

// 128-bit shared-memory accesses with explicit 32-bit addresses (keeps the address arithmetic of the
// pass loop to one shift + add per neighbour)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4 &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace mt_snip_dev
