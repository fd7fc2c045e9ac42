// K2b -- SNIP background kernel for aligned volumes: contiguous channel runs, persistent CTAs, bulk-copy prefetch.
//
// Same arithmetic as mt_snip.cu (bkg.py:53-114; results are bit-identical), different organisation.  Measured on the
// strided kernel (B200, 32 768 px x 4096 ch, 14 passes): 0.45 ms for prologue + epilogue -- latency bound, the two
// resident CTAs of an SM wait for HBM and L2 with nothing else to run -- plus 0.046 ms per pass at 91 % of the
// shared-memory rate (12 bytes per channel, pass and pixel).  This kernel attacks both:
//   * a thread owns R CONSECUTIVE channels of the CTA's 4 pixels (registers) and the shared copy is swizzled on
//     16-byte records (slot = i ^ ((i >> 3) & 7)) so that runs, neighbours at any uniform offset and the table-driven
//     gathers are all free of bank conflicts.  In the late passes the window offsets (a, b) are <= (4, 3) and uniform
//     over a run: neighbours then come from the thread's own registers and only the a + b halo records are read from
//     shared memory (descending order, in place) -- 8 LDS + 8 STS per run become <= 7 LDS + 8 STS;
//   * CTAs are persistent (2 per SM).  When the last pass has been read, the shared buffer is dead: one thread
//     issues cp.async.bulk copies of the NEXT quad's rows into it (mbarrier completion) and the epilogue (exp,
//     residual, 128-bit stores) runs while they land, so the prologue finds its counts in shared memory;
//   * loads and stores of a run are 128-bit (8 fp16 = one STG.128 per pixel and plane instead of 16 STG.16).
// Preconditions (checked by the host, which otherwise launches the strided kernel): 16-byte aligned rows and crops,
// n_ch % 8 == 0, n_ch <= 4096, boxcar == 5.
#include <cuda_fp16.h>
#include <cstdlib>
#include <type_traits>

#include "mt_common.cuh"
#include "mt_f32x2.cuh"
#include "mt_ptx.cuh"
#include "mt_snip_dev.cuh"

namespace {

using namespace mt_snip_dev;
using namespace mt::ptx;

constexpr int kRunThreads = 512;
constexpr int kRunPix = 4;
constexpr uint32_t kDescGeneral = 0u;  // neighbours through the gather table
constexpr uint32_t kDescSkip = 1u;     // run past the end of the window
constexpr uint32_t kDescFast = 0x100u; // | a << 4 | b: uniform offsets, neighbours from registers + halo
constexpr uint32_t kDescKind = 0x1ffu; // bits 16..23 of a descriptor: the records of the run that somebody reads through
                                       // shared memory in this pass (they are published after the previous pass)

__host__ __device__ constexpr int run_len(int n_ch) { return n_ch <= 1024 ? 2 : (n_ch <= 2048 ? 4 : 8); }

__host__ __device__ inline bool fast_pair(int a, int b) {
    return a >= 1 && a <= 4 && (b == a - 1 || (b == a && a <= 3));
}

// position of the gather word of channel j of run t inside a table row
__host__ __device__ inline int word_pos(int R, int t, int j) {
    return R == 8 ? (j >> 2) * (kRunThreads * 4) + t * 4 + (j & 3) : t * R + j;
}

// swizzled byte offset of record i
__host__ __device__ inline uint32_t rec_off(int i) { return (uint32_t)(i ^ ((i >> 3) & 7)) << 4; }

// ---- tables ------------------------------------------------------------------------------------------------------
// out: n_pass rows of ld words: [0, 512 R) gather words (swizzled byte offsets lo | hi << 16; entries past n_ch point
// at their own record), [512 R, 512 R + 512) one descriptor per run.  The kernel copies one row per pass into shared
// memory (cp.async.bulk) and every thread reads its R words from there with 128-bit loads: the word of (run t,
// channel j) sits at word_pos(t, j) -- for R = 8 in two planes of four words per run, which keeps the loads at a lane
// stride of 16 bytes (no bank conflicts).
__global__ void snip_runs_tables_kernel(const uint32_t *__restrict__ lohi, int n_pass, int n_ch, int R,
                                        uint32_t *__restrict__ out, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, pass = blockIdx.y;
    const int n_tab = kRunThreads * R;
    uint32_t *row = out + (long long)pass * ld;
    if (i < n_tab) {
        uint32_t w = rec_off(i) | (rec_off(i) << 16);
        if (pass < n_pass && i < n_ch) {
            const uint32_t v = lohi[(long long)pass * n_ch + i];
            w = rec_off((int)(v & 0xffffu)) | (rec_off((int)(v >> 16)) << 16);
        }
        row[word_pos(R, i / R, i % R)] = w;
    }
    if (i < kRunThreads) {
        const int i0 = i * R;
        uint32_t d = kDescSkip;
        if (pass < n_pass && i0 < n_ch) {
            d = kDescGeneral;
            const uint32_t v0 = lohi[(long long)pass * n_ch + i0];
            const int a = i0 - (int)(v0 & 0xffffu), b = (int)(v0 >> 16) - i0;
            bool uniform = i0 + R <= n_ch;
            for (int j = 1; j < R && uniform; ++j) {
                const uint32_t v = lohi[(long long)pass * n_ch + i0 + j];
                uniform = (i0 + j - (int)(v & 0xffffu)) == a && ((int)(v >> 16) - i0 - j) == b;
            }
            if (uniform && fast_pair(a, b)) d = kDescFast | (uint32_t)(a << 4) | (uint32_t)b;
        }
        row[n_tab + i] = d;
    }
}

// second step: publish masks.  Record lo_i / hi_i of channel i is read through shared memory unless the reading run
// takes it from its own registers (register path and the record inside the run).
__global__ void snip_runs_masks_kernel(const uint32_t *__restrict__ lohi, int n_pass, int n_ch, int R, uint32_t *out, int ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, pass = blockIdx.y;
    if (i >= n_ch || pass >= n_pass) return;
    uint32_t *desc = out + (long long)pass * ld + kRunThreads * R;
    const int i0 = (i / R) * R;
    const bool fast = (desc[i / R] & kDescFast) != 0;  // kind bits are final (first kernel); only mask bits change here
    const uint32_t v = lohi[(long long)pass * n_ch + i];
    const int nb[2] = {(int)(v & 0xffffu), (int)(v >> 16)};
    for (int s = 0; s < 2; ++s) {
        const int r = nb[s];
        if (fast && r >= i0 && r < i0 + R) continue;
        atomicOr(desc + r / R, 0x10000u << (r % R));
    }
}

// ---- vector access to a run --------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void load_run(const float *p, float (&v)[R]) {
    if constexpr (R == 2) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        v[0] = t.x, v[1] = t.y;
    } else {
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4 *>(p + 4 * q);
            v[4 * q] = t.x, v[4 * q + 1] = t.y, v[4 * q + 2] = t.z, v[4 * q + 3] = t.w;
        }
    }
}
// a run of the staged rows in shared memory.  Eight fp32 channels are two 16-byte halves at a lane stride of 32 bytes:
// lanes 4..7 of every group of eight fetch the upper half first, which makes both loads free of bank conflicts.
template <int R>
__device__ __forceinline__ void load_run_staged(const float *p, float (&v)[R], int swap) {
    if constexpr (R == 8) {
        const float4 a = *reinterpret_cast<const float4 *>(p + (swap ? 4 : 0));
        const float4 b = *reinterpret_cast<const float4 *>(p + (swap ? 0 : 4));
        const float4 lo = swap ? b : a, hi = swap ? a : b;
        v[0] = lo.x, v[1] = lo.y, v[2] = lo.z, v[3] = lo.w, v[4] = hi.x, v[5] = hi.y, v[6] = hi.z, v[7] = hi.w;
    } else {
        load_run<R>(p, v);
    }
}
template <int R>
__device__ __forceinline__ void load_run(const __half *p, float (&v)[R]);
template <int R>
__device__ __forceinline__ void load_run_staged(const __half *p, float (&v)[R], int) {
    load_run<R>(p, v);
}
template <int R>
__device__ __forceinline__ void load_run(const __half *p, float (&v)[R]) {
    uint32_t w[R / 2];
    if constexpr (R == 2) {
        w[0] = *reinterpret_cast<const uint32_t *>(p);
    } else if constexpr (R == 4) {
        const uint2 t = *reinterpret_cast<const uint2 *>(p);
        w[0] = t.x, w[1] = t.y;
    } else {
        const uint4 t = *reinterpret_cast<const uint4 *>(p);
        w[0] = t.x, w[1] = t.y, w[2] = t.z, w[3] = t.w;
    }
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
        v[2 * q] = f.x, v[2 * q + 1] = f.y;
    }
}
// two consecutive channels (boxcar halo); p is 2-element aligned
__device__ __forceinline__ void load_pair(const float *p, float &a, float &b) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    a = t.x, b = t.y;
}
__device__ __forceinline__ void load_pair(const __half *p, float &a, float &b) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2 *>(p));
    a = t.x, b = t.y;
}

template <int R>
__device__ __forceinline__ void store_run(float *p, const float (&v)[R]) {
    if constexpr (R == 2) {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int q = 0; q < R / 4; ++q)
            *reinterpret_cast<float4 *>(p + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}
template <int R>
__device__ __forceinline__ void store_run(__half *p, const __half (&v)[R]) {
    uint32_t w[R / 2];
#pragma unroll
    for (int q = 0; q < R / 2; ++q) {
        const __half2 h = __halves2half2(v[2 * q], v[2 * q + 1]);
        w[q] = *reinterpret_cast<const uint32_t *>(&h);
    }
    if constexpr (R == 2)
        *reinterpret_cast<uint32_t *>(p) = w[0];
    else if constexpr (R == 4)
        *reinterpret_cast<uint2 *>(p) = make_uint2(w[0], w[1]);
    else
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

template <int P>
__device__ __forceinline__ float &comp(float4 &v) {
    if constexpr (P == 0) return v.x;
    if constexpr (P == 1) return v.y;
    if constexpr (P == 2) return v.z;
    return v.w;
}

// bounded spin: a protocol bug or a faulting copy must surface as a launch failure, not hang the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 0xFFFu) == 0xFFFu) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000ll) __trap();  // (no printf: its stack frame costs the kernel registers)
        }
    }
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- one clipping pass of a run with uniform offsets (A below, B above): descending order, in place -------------
template <int R, int A, int B>
__device__ __forceinline__ void clip_run(float4 (&cur)[R], uint32_t sb_base, int i0) {
    float4 old[R];
#pragma unroll
    for (int j = 0; j < R; ++j) old[j] = cur[j];
#pragma unroll
    for (int j = R - 1; j >= 0; --j) {
        float4 lo, hi;
        if (j - A >= 0)
            lo = old[j - A >= 0 ? j - A : 0];
        else
            lo = lds128(sb_base + rec_off(i0 + j - A));
        if (j + B < R)
            hi = old[j + B < R ? j + B : 0];
        else
            hi = lds128(sb_base + rec_off(i0 + j + B));
        clip4(lo, hi, cur[j]);
    }
}

template <typename T, int R, bool FAST>
__global__ void __launch_bounds__(kRunThreads, 2)
snip_runs_kernel(const T *__restrict__ spec, long long n_px, long long ld_spec, int ch0, int n_ch,
                 const uint32_t *__restrict__ tabs, int ld_tab, int n_pass, void *__restrict__ out_, long long ld_out,
                 int out_mode, int *__restrict__ overflow) {
    // [512 R] swizzled records x (4 pixels); between quads: the next quad's raw rows.  Behind them two table-row
    // buffers and three mbarriers (no static shared memory: the record addresses are formed with OR and need the
    // 128-byte aligned base)
    extern __shared__ __align__(128) float4 sb[];
    const int tid = threadIdx.x;
    const int i0 = tid * R;
    const bool active = i0 < n_ch;
    const uint32_t sb_base = smem_u32(sb);
    if (sb_base & 127u) __trap();
    const long long n_quad = (n_px + kRunPix - 1) / kRunPix;
    const uint32_t row_bytes = (uint32_t)n_ch * (uint32_t)sizeof(T);
    const T *raw = reinterpret_cast<const T *>(sb);
    // own records: slot = (i0 & ~7) + (j ^ sx)
    const uint32_t own_base = sb_base + ((uint32_t)(i0 & ~7) << 4);
    const uint32_t sx4 = (uint32_t)((i0 & 7) ^ ((i0 >> 3) & 7)) << 4;
    const float wgt = 1.0f / 5.0f;
    // table rows: two shared buffers behind the records, filled one pass ahead by bulk copies
    const uint32_t tab_bytes = (uint32_t)ld_tab * 4u;
    const uint32_t tab_base = sb_base + (uint32_t)sizeof(float4) * kRunThreads * R;
    const uint32_t bar = tab_base + 2u * tab_bytes;  // [0]: staged rows, [1], [2]: table buffers
    const uint32_t my_desc = (uint32_t)(kRunThreads * R + tid) * 4u;
    const uint32_t my_words = (uint32_t)(R == 8 ? tid * 4 : tid * R) * 4u;

    auto issue = [&](long long quad) {
        const long long px0 = quad * kRunPix;
        const int rows = (int)(n_px - px0 < kRunPix ? n_px - px0 : kRunPix);
        mbar_arrive_expect_tx(bar, row_bytes * (uint32_t)rows);
        for (int p = 0; p < rows; ++p)
            bulk_load(sb_base + (uint32_t)p * row_bytes, spec + (px0 + p) * ld_spec + ch0, row_bytes, bar);
    };
    // g counts the passes this CTA has started; pass g uses buffer g & 1 and its barrier's phase (g >> 1) & 1
    auto issue_table = [&](uint32_t g, int row) {
        const uint32_t b = g & 1u;
        mbar_arrive_expect_tx(bar + 8u + 8u * b, tab_bytes);
        bulk_load(tab_base + b * tab_bytes, tabs + (long long)row * ld_tab, tab_bytes, bar + 8u + 8u * b);
    };
    auto wait_table = [&](uint32_t g) {
        mbar_wait_bounded(bar + 8u + 8u * (g & 1u), (g >> 1) & 1u);
        return tab_base + (g & 1u) * tab_bytes;
    };

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 8u, 1);
        mbar_init(bar + 16u, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0 && (long long)blockIdx.x < n_quad) {
        issue(blockIdx.x);
        if (n_pass > 0) issue_table(0u, 0);
    }

    uint32_t g = 0;
    uint32_t parity = 0;
    for (long long quad = blockIdx.x; quad < n_quad; quad += gridDim.x, parity ^= 1u) {
        const long long px0 = quad * kRunPix;
        float4 cur[R];
        mbar_wait_bounded(bar, parity);

        // ---- boxcar mean (zero padded, bkg.py:93-97) and log1p (bkg.py:100-101) from the staged rows --------------
        // (every thread takes part: the halo channels travel by warp shuffle; runs past the window compute values that
        // are never stored, from addresses inside the buffer)
        {
            auto one_pixel = [&](auto pc) {
                constexpr int P = decltype(pc)::value;
                const T *row = raw + (long long)P * n_ch + i0;
                float v[R + 4];
                float own[R];
                load_run_staged<R>(row, own, (tid >> 2) & 1);
                v[0] = __shfl_up_sync(0xffffffffu, own[R - 2], 1);
                v[1] = __shfl_up_sync(0xffffffffu, own[R - 1], 1);
                v[R + 2] = __shfl_down_sync(0xffffffffu, own[0], 1);
                v[R + 3] = __shfl_down_sync(0xffffffffu, own[1], 1);
                if ((tid & 31) == 0) {
                    v[0] = v[1] = 0.f;
                    if (i0 >= 2) load_pair(row - 2, v[0], v[1]);
                }
                if (i0 + R + 2 > n_ch) {
                    v[R + 2] = v[R + 3] = 0.f;
                } else if ((tid & 31) == 31) {
                    load_pair(row + R, v[R + 2], v[R + 3]);
                }
#pragma unroll
                for (int j = 0; j < R; ++j) v[2 + j] = own[j];
                // the strided kernel's arithmetic, bit for bit: fused multiply-add chains away from the edges of the
                // window (its packed mul + add pairs compile to FFMA2), separately rounded products in the two channels
                // next to each edge
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    float acc = 0.f;
                    const bool edge = (i0 + j < 2) || (i0 + j + 2 >= n_ch);
                    if (!edge) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) acc = __fmaf_rn(v[j + k], wgt, acc);
                    } else {
#pragma unroll
                        for (int k = 0; k < 5; ++k) acc = __fadd_rn(acc, __fmul_rn(v[j + k], wgt));
                    }
                    comp<P>(cur[j]) = finite_or_zero(log_1p<FAST>(acc));
                }
            };
            one_pixel(std::integral_constant<int, 0>{});
            one_pixel(std::integral_constant<int, 1>{});
            one_pixel(std::integral_constant<int, 2>{});
            one_pixel(std::integral_constant<int, 3>{});
        }
        __syncthreads();  // every raw value has been read: the buffer becomes the log-background
        // only the records somebody reads through shared memory in the coming pass are published
        auto publish = [&](uint32_t tb) {
            uint32_t dd;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dd) : "r"(tb + my_desc));
            const uint32_t m = (dd >> 16) & 0xffu;
            if (m == (1u << R) - 1u) {
#pragma unroll
                for (int j = 0; j < R; ++j) sts128(own_base | (((uint32_t)j << 4) ^ sx4), cur[j]);
            } else if (m) {
#pragma unroll
                for (int j = 0; j < R; ++j)
                    if (m & (1u << j)) sts128(own_base | (((uint32_t)j << 4) ^ sx4), cur[j]);
            }
        };
        if (n_pass > 0) publish(wait_table(g));
        __syncthreads();

        // ---- clipping passes (bkg.py:53-66, 104-112), Jacobi style ------------------------------------------------
        const bool more_quads = quad + gridDim.x < n_quad;
        for (int pass = 0; pass < n_pass; ++pass, ++g) {
            const bool last = pass + 1 == n_pass;
            // the other buffer was last read before the barriers behind us: fetch the coming pass's row into it
            if (tid == 0 && (!last || more_quads)) issue_table(g + 1u, last ? 0 : pass + 1);
            const uint32_t tb = tab_base + (g & 1u) * tab_bytes;  // this pass's row (awaited before its records were published)
            uint32_t d;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(d) : "r"(tb + my_desc));
            d &= kDescKind;
            if (d == kDescGeneral) {
                constexpr int W = R >= 4 ? 4 : R;
#pragma unroll
                for (int q = 0; q < R / W; ++q) {
                    uint32_t w[W];
                    if constexpr (W == 4)
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                     : "r"(tb + my_words + (uint32_t)q * (kRunThreads * 16u)));
                    else
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(tb + my_words));
#pragma unroll
                    for (int k = 0; k < W; ++k) {
                        const float4 a = lds128(sb_base + (w[k] & 0xffffu));
                        const float4 b = lds128(sb_base + (w[k] >> 16));
                        clip4(a, b, cur[q * W + k]);
                    }
                }
            } else if (d != kDescSkip) {
                switch (d & 0xffu) {
                    case 0x10: clip_run<R, 1, 0>(cur, sb_base, i0); break;
                    case 0x11: clip_run<R, 1, 1>(cur, sb_base, i0); break;
                    case 0x21: clip_run<R, 2, 1>(cur, sb_base, i0); break;
                    case 0x22: clip_run<R, 2, 2>(cur, sb_base, i0); break;
                    case 0x32: clip_run<R, 3, 2>(cur, sb_base, i0); break;
                    case 0x33: clip_run<R, 3, 3>(cur, sb_base, i0); break;
                    default: clip_run<R, 4, 3>(cur, sb_base, i0); break;
                }
            }
            if (last) {  // the result of the last pass stays in registers
                ++g;
                break;
            }
            __syncthreads();
            publish(wait_table(g + 1u));
            __syncthreads();
        }

        // ---- the shared buffer is dead: fetch the next quad's rows behind the epilogue ----------------------------
        __syncthreads();
        if (tid == 0 && quad + gridDim.x < n_quad) {
            fence_proxy_async();
            issue(quad + gridDim.x);
        }

        // ---- back to counts (bkg.py:113-114) and output -----------------------------------------------------------
        if (active) {
            const bool need_y = out_mode != MT_SNIP_BACKGROUND;
            auto one_pixel = [&](auto pc) {
                constexpr int P = decltype(pc)::value;
                const long long px = px0 + P;
                if (px >= n_px) return;
                // y[] ends up holding what is emitted: the background, or the residual y - bkg
                float y[R];
#pragma unroll
                for (int j = 0; j < R; ++j) y[j] = 0.f;
                if (need_y) load_run<R>(spec + px * ld_spec + ch0 + i0, y);
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float bkg = finite_or_zero(exp_m1<FAST>(comp<P>(cur[j])));
                    y[j] = need_y ? __fsub_rn(y[j], bkg) : bkg;
                }
                if (out_mode == MT_SNIP_RESIDUAL_HILO_F16) {
                    __half hi[R];
                    bool big = false;
#pragma unroll
                    for (int j = 0; j < R; ++j) {
                        hi[j] = __float2half_rn(y[j]);
                        big |= !(fabsf(y[j]) <= 65504.f);
                    }
                    __half *hrow = static_cast<__half *>(out_) + px * ld_out + i0;
                    store_run<R>(hrow, hi);
#pragma unroll
                    for (int j = 0; j < R; ++j) hi[j] = __float2half_rn(__fsub_rn(y[j], __half2float(hi[j])));
                    store_run<R>(hrow + n_ch, hi);
                    if (big) atomicOr(overflow, 1);
                } else {
                    float *orow = static_cast<float *>(out_) + px * ld_out + i0;
                    if (out_mode == MT_SNIP_RESIDUAL_HILO) {
                        float hi[R];
#pragma unroll
                        for (int j = 0; j < R; ++j) {
                            hi[j] = mt::tf32_round(y[j]);
                            y[j] = __fsub_rn(y[j], hi[j]);
                        }
                        store_run<R>(orow, hi);
                        orow += n_ch;
                    }
                    store_run<R>(orow, y);
                }
            };
            one_pixel(std::integral_constant<int, 0>{});
            one_pixel(std::integral_constant<int, 1>{});
            one_pixel(std::integral_constant<int, 2>{});
            one_pixel(std::integral_constant<int, 3>{});
        }
    }
}

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return v ? atoi(v) : fallback;
}

template <typename T, int R, bool FAST>
int launch_runs(const void *spec, int64_t n_px, int64_t ld_spec, int ch0, int n_ch, const uint32_t *tabs, int ld_tab,
                int n_pass, void *out, int64_t ld_out, int out_mode, int *overflow, cudaStream_t stream) {
    auto kern = snip_runs_kernel<T, R, FAST>;
    const size_t smem = sizeof(float4) * (size_t)kRunThreads * R + 2 * sizeof(uint32_t) * (size_t)ld_tab + 32;  // records, two table rows, barriers
    static int smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 16) smem_set[dev] = (int)smem;
    }
    const long long n_quad = (n_px + kRunPix - 1) / kRunPix;
    static const int per_sm = env_int("MT_SNIP_RUNS_CTAS_PER_SM", 2);
    const long long slots = (long long)mt::num_sms() * per_sm;
    const unsigned grid = (unsigned)(n_quad < slots ? n_quad : slots);
    kern<<<grid, kRunThreads, smem, stream>>>(static_cast<const T *>(spec), n_px, ld_spec, ch0, n_ch, tabs, ld_tab, n_pass,
                                              out, ld_out, out_mode, overflow);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

template <typename T>
int dispatch_runs(const void *spec, int64_t n_px, int64_t ld_spec, int ch0, int n_ch, const uint32_t *tabs, int ld_tab,
                  int n_pass, void *out, int64_t ld_out, int out_mode, int *overflow, cudaStream_t stream) {
    static const bool fast = env_int("MT_SNIP_FASTMATH", 1) != 0;
#define MT_RUNS_CASE(R)                                                                                                \
    if (run_len(n_ch) == R) {                                                                                          \
        if (fast)                                                                                                      \
            return launch_runs<T, R, true>(spec, n_px, ld_spec, ch0, n_ch, tabs, ld_tab, n_pass, out, ld_out, out_mode, \
                                           overflow, stream);                                                          \
        return launch_runs<T, R, false>(spec, n_px, ld_spec, ch0, n_ch, tabs, ld_tab, n_pass, out, ld_out, out_mode,    \
                                        overflow, stream);                                                             \
    }
    MT_RUNS_CASE(2)
    MT_RUNS_CASE(4)
    MT_RUNS_CASE(8)
#undef MT_RUNS_CASE
    return mt::fail(MT_ERR_UNSUPPORTED, "mt_snip_runs: n_ch=%d", n_ch);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int mt_snip_runs_pitch(int32_t n_ch) {
    if (n_ch <= 0 || n_ch > 4096 || (n_ch & 7)) return 0;
    return kRunThreads * run_len(n_ch) + kRunThreads;
}

extern "C" int mt_snip_runs_tables(const uint32_t *lohi_dev, int32_t n_pass, int32_t n_ch, uint32_t *tabs_dev,
                                   int32_t ld_tabs, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(lohi_dev && tabs_dev, "mt_snip_runs_tables: null pointer");
    MT_REQUIRE(n_pass >= 0 && n_ch > 0, "mt_snip_runs_tables: bad sizes");
    MT_REQUIRE(mt_snip_runs_pitch(n_ch) > 0 && ld_tabs == mt_snip_runs_pitch(n_ch),
               "mt_snip_runs_tables: n_ch=%d needs a row pitch of %d words (0 = not available), got %d", n_ch,
               mt_snip_runs_pitch(n_ch), ld_tabs);
    const int R = run_len(n_ch);
    dim3 grid((kRunThreads * R + 255) / 256, n_pass);
    snip_runs_tables_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(lohi_dev, n_pass, n_ch, R, tabs_dev,
                                                                                  ld_tabs);
    MT_CUDA_TRY(cudaGetLastError());
    if (n_pass > 0) {
        dim3 grid2((n_ch + 255) / 256, n_pass);
        snip_runs_masks_kernel<<<grid2, 256, 0, static_cast<cudaStream_t>(stream_)>>>(lohi_dev, n_pass, n_ch, R, tabs_dev,
                                                                                      ld_tabs);
        MT_CUDA_TRY(cudaGetLastError());
    }
    return MT_OK;
}

extern "C" int mt_snip_runs_eligible(const void *spec_dev, int32_t dtype, int64_t ld_spec, int32_t ch0, int32_t n_ch,
                                     int32_t boxcar, const void *out_dev, int64_t ld_out, int32_t out_mode) {
    static const bool off = env_int("MT_SNIP_RUNS", 1) == 0;
    if (off || mt_snip_runs_pitch(n_ch) == 0 || boxcar != 5) return 0;
    const int64_t es = dtype == MT_DTYPE_F16 ? 2 : (dtype == MT_DTYPE_F32 ? 4 : 0);
    if (es == 0) return 0;
    if (!aligned16(spec_dev) || (ld_spec * es) % 16 || ((int64_t)ch0 * es) % 16) return 0;
    const int64_t os = out_mode == MT_SNIP_RESIDUAL_HILO_F16 ? 2 : 4;
    if (!aligned16(out_dev) || (ld_out * os) % 16) return 0;
    return 1;
}

extern "C" int mt_snip_runs(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec, int32_t ch0,
                            int32_t n_ch, const uint32_t *tabs_dev, int32_t ld_tabs, int32_t n_pass, int32_t boxcar,
                            void *out_dev, int64_t ld_out, int32_t out_mode, int32_t *overflow_dev, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && tabs_dev && out_dev, "mt_snip_runs: null pointer");
    MT_REQUIRE(n_px >= 0 && n_ch > 0 && ch0 >= 0 && ld_spec >= (int64_t)ch0 + n_ch,
               "mt_snip_runs: bad sizes n_px=%lld n_ch=%d ch0=%d ld=%lld", (long long)n_px, n_ch, ch0, (long long)ld_spec);
    MT_REQUIRE(n_pass >= 0, "mt_snip_runs: n_pass=%d", n_pass);
    MT_REQUIRE(out_mode >= MT_SNIP_BACKGROUND && out_mode <= MT_SNIP_RESIDUAL_HILO_F16, "mt_snip_runs: out_mode=%d", out_mode);
    MT_REQUIRE(out_mode != MT_SNIP_RESIDUAL_HILO_F16 || overflow_dev, "mt_snip_runs: fp16 planes need an overflow flag");
    MT_REQUIRE(ld_out >= (out_mode >= MT_SNIP_RESIDUAL_HILO ? 2 : 1) * (int64_t)n_ch, "mt_snip_runs: ld_out=%lld too small",
               (long long)ld_out);
    MT_REQUIRE(ld_tabs == mt_snip_runs_pitch(n_ch) && ld_tabs > 0, "mt_snip_runs: table pitch %d for n_ch=%d", ld_tabs, n_ch);
    if (!mt_snip_runs_eligible(spec_dev, dtype, ld_spec, ch0, n_ch, boxcar, out_dev, ld_out, out_mode))
        return mt::fail(MT_ERR_UNSUPPORTED,
                        "mt_snip_runs: needs 16-byte aligned rows and crops, n_ch %% 8 == 0, n_ch <= 4096, boxcar == 5 "
                        "(mt_snip_runs_eligible); use mt_snip / mt_snip_scaled otherwise");
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    {
        static const int cap = env_int("MT_SNIP_DEBUG_NPASS", -1);  // measurement knob, see mt_snip.cu
        if (cap >= 0 && cap < n_pass) n_pass = cap;
    }
    if (dtype == MT_DTYPE_F32)
        return dispatch_runs<float>(spec_dev, n_px, ld_spec, ch0, n_ch, tabs_dev, ld_tabs, n_pass, out_dev, ld_out, out_mode,
                                    overflow_dev, stream);
    return dispatch_runs<__half>(spec_dev, n_px, ld_spec, ch0, n_ch, tabs_dev, ld_tabs, n_pass, out_dev, ld_out, out_mode,
                                 overflow_dev, stream);
}
