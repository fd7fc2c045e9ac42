// K3 -- streaming per-pixel amplitude solve on the 5th-generation tensor cores.
//
// For fixed global parameters the MAPS model is linear in the element amplitudes and the Gram
// matrix is pixel independent, so the reference's per-pixel Adam loop
// (opt.py:249-317) collapses to X[P,E] = Y[P,C] . W[C,E] with W the pseudo-inverse of the
// element basis.  Each spectrum byte is read from HBM exactly once:
//
//   warp 0   TMA producer : 2-D bulk-tensor loads of a [256 px x 128 B] slab of spectra
//                           (evict-first) and the matching [N x 128 B] slab of W (evict-last,
//                           L2 resident) into a ring of 128B-swizzled shared-memory stages
//   warp 1   MMA issuer   : tcgen05.mma cta_group::1, M=128 (x2 pixel sub-tiles sharing the W
//                           slab), N = n_planes*e_pad, kind::f16 for fp16 volumes / kind::tf32 for
//                           fp32; accumulators live in TMEM, double buffered across pixel tiles
//   warps 2-5 epilogue    : tcgen05.ld the finished tile, recombine the digit planes of W
//                           (Horner, fp32), scale, store element-major amplitudes x[e][p]
//                           (coalesced over pixels)
//
// W is stored as n_planes small-integer "digit" planes stacked along N (10 bits each, exact in
// fp16/tf32).  Spectra are raw counts (exact below 2048), so every product and every partial sum
// is an integer: as long as sum_k y_k*|digit_k| < 2^24 the fp32 accumulation in the tensor core
// is EXACT, and the only rounding left is the fp32 Horner recombination (~1e-7 relative).  With
// full-mantissa operands the tensor-core accumulator (which truncates) costs ~4e-6 of the pixel
// maximum instead -- measured, see DESIGN.md.
// Bound: HBM (C*sizeof(in) + E*4 bytes per pixel); tensor-pipe use stays well below peak.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>

#include "mt_common.cuh"
#include "mt_ptx.cuh"

namespace {

using namespace mt::ptx;

constexpr int kFitThreads = 192;       // 6 warps: producer, MMA, 4 epilogue
constexpr int kTileRows = 128;         // UMMA M
constexpr int kRowBytes = 128;         // one swizzle-128B row = one k-block per row
constexpr int kSubTileBytes = kTileRows * kRowBytes;  // 16 KiB
constexpr int kSmemLimit = 227 * 1024;

struct FitParams {
    long long n_px;
    long long ld_x;
    long long px_offset;  // added to local pixel numbers written to neg_list
    int *neg_list;        // optional: pixels whose solution has a negative component
    int *neg_count;
    int pixel_major;      // x layout: 0 = x[e*ld_x + p], 1 = x[p*ld_x + e]
    const float *hinv32;  // optional G^-1 (fp32 [n_comp][n_comp]): non-negative fix-up in the epilogue
    int reg_path;         // epilogue keeps the whole pixel in registers (e_pad <= 32, element-major)
    MtPeers peers;        // extra destinations on other GPUs (n_peers == 0 and multicast == NULL: none)
    float *x;
    const float *colscale;
    int n_tiles;
    // tile schedule: every CTA takes k_full whole tiles (tile = block + k * grid), then ONE partial tile of the remainder
    // region [rem_row0, n_px): rem_base (+1 for the first rem_extra CTAs) granules of 64 pixels, contiguous per CTA
    int k_full, rem_base, rem_extra;
    long long rem_row0;
    uint32_t idesc64;  // instruction descriptor of the 64-row MMA of a partial tile's last sub-tile
    int bulk;          // epilogue stores through shared memory + cp.async.bulk (reg path without peers)
    int debug;         // measurement knobs (MT_FIT_DEBUG bit mask): 1 no stores, 2 no pack loads, 4 no MMAs, 8 no fix-up
    int kb_begin, kb_end;
    int bke;  // elements per k-block (64 for fp16, 32 for fp32)
    int n_mma, e_pad, n_comp, n_planes;
    int n_stages;
    uint32_t stage_bytes;
    uint32_t tmem_cols;
    uint32_t idesc;
    float plane_ratio;
};

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a protocol bug must surface as a launch failure, not hang the GPU
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 0xFFFu) == 0xFFFu) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000ll) {
                printf("mt_fit_contract: barrier wait timed out (block %d thread %d)\n", blockIdx.x,
                       threadIdx.x);
                __trap();
            }
        }
    }
}

// Non-negative fix-up of one pixel inside the epilogue, registers only: block principal pivoting
// restricted to zero sets of at most two components (closed-form 1x1 / 2x2 Schur complements on
// H = G^-1, SURVEY.md 0.4 / csrc/mt_nnls.cu).  Returns false when the pixel needs a larger zero set
// or does not settle within a few pivots; the caller then leaves the unconstrained values in place
// and lists the pixel for the general solver (mt_nnls_list).
__device__ __forceinline__ bool fixup_small(float (&x)[32], int E, const float *hs) {
    float xu[32];
    unsigned A = 0u;
    float xmax = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        xu[e] = x[e];
        if (e < E) {
            A |= (xu[e] < 0.f) ? (1u << e) : 0u;
            xmax = fmaxf(xmax, fabsf(xu[e]));
        }
    }
    const float tol = 1e-7f * xmax;
    bool settled = false;
    for (int iter = 0; iter < 6; ++iter) {
        const int n = __popc(A);
        if (n > 2) break;
        if (n == 0) {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = xu[e];
            unsigned V = 0u;
#pragma unroll
            for (int e = 0; e < 32; ++e)
                if (e < E && xu[e] < -tol) V |= 1u << e;
            if (V == 0u) {
                settled = true;
                break;
            }
            A = V;
            continue;
        }
        const int a = __ffs(A) - 1, b = 31 - __clz(A);
        float xa = 0.f, xb = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            xa = (e == a) ? xu[e] : xa;
            xb = (e == b) ? xu[e] : xb;
        }
        // strongly overlapping components make H_AA nearly singular (z large, corrections cancel):
        // the small solve and the update run in fp64, H itself is kept in fp32
        const double haa = hs[a * E + a], hbb = hs[b * E + b], hab = hs[a * E + b];
        double za, zb;
        if (n == 1) {
            za = (double)xa / haa;
            zb = 0.0;
        } else {
            const double det = haa * hbb - hab * hab;
            za = (hbb * xa - hab * xb) / det;
            zb = (haa * xb - hab * xa) / det;
        }
        // duals on the zero set are -z (in units of x / H); primal on the free set
        const double tol_y = (double)tol / fmin(haa, hbb);
        unsigned V = 0u;
        if (-za < -tol_y) V |= 1u << a;
        if (n == 2 && -zb < -tol_y) V |= 1u << b;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            if (e < E) {
                if ((A >> e) & 1u) {
                    x[e] = 0.f;
                } else {
                    const float v = (float)((double)xu[e] - (double)hs[e * E + a] * za - (double)hs[e * E + b] * zb);
                    x[e] = v;
                    if (v < -tol) V |= 1u << e;
                }
            }
        }
        if (V == 0u) {
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = fmaxf(x[e], 0.f);
            settled = true;
            break;
        }
        A ^= V;
    }
    if (!settled) {
        // hand the UNCONSTRAINED solution to the general solver
#pragma unroll
        for (int e = 0; e < 32; ++e) x[e] = xu[e];
    }
    return settled;
}

// CL > 1: clusters of CL CTAs share every slab of the pseudo-inverse pack -- CTA r of a cluster loads rows
// [r, r + 1) * N / CL of it and the TMA multicasts them into the same stage of all CL rings, so the pack crosses the L2
// once per cluster and k-block instead of once per CTA.  Measured why: the streaming rate follows
// 7.8 TB/s / (1 + pack slab / spectra slab) -- 6.5 TB/s at N = 48, 6.1 at N = 72, 5.6-5.8 at N = 96 -- i.e. the L2's
// aggregate delivery rate bounds the kernel, not HBM, and a third of what it delivered at N = 96 was the pack.
// The CTAs of a cluster walk their tiles in lockstep (stage by stage: a stage is refilled when ALL of them have
// consumed it -- the MMA warps commit to every CTA's empty barrier); a CTA with one tile less runs a round without
// spectra so that its partners still get their slabs.
template <int KIND, int M_TILES, int CL>
__global__ void __launch_bounds__(kFitThreads, 1)
fit_contract_kernel(const __grid_constant__ CUtensorMap tm_spec, const __grid_constant__ CUtensorMap tm_spec64,
                    const __grid_constant__ CUtensorMap tm_w, const FitParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t ring = (raw_addr + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-B alignment
    uint8_t *after_ring = smem_raw + (ring - raw_addr) + (size_t)p.n_stages * p.stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(after_ring);
    const uint32_t bar0 = smem_u32(bars);
    // barrier slots: full[s], empty[s], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(p.n_stages + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * p.n_stages + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (uint32_t)(2 * p.n_stages + 2 + a); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * p.n_stages + 4);
    float *cs = reinterpret_cast<float *>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int block_rows = kTileRows * M_TILES;
    const uint32_t acc_cols = (uint32_t)(M_TILES * p.n_mma);
    // The k-th tile of this CTA: first pixel and size in 64-pixel granules (0: no such tile).  Whole tiles are dealt
    // round-robin; what is left of the map (less than one tile per CTA) is cut into 64-pixel granules and every CTA
    // takes its share as one last, partial tile -- a 128-row MMA per full pair of granules, a 64-row MMA for an odd
    // one -- so that no SM idles through a last wave (c2: 512 half-tiles on 148 SMs ran as 4 rounds for 3.46 of work).
    constexpr int kGranPerTile = 2 * M_TILES;
    auto tile_of = [&](int b, int k, long long &row0) -> int {
        if (k < p.k_full) {
            const long long tile = (long long)b + (long long)k * gridDim.x;
            row0 = tile * block_rows;
            return tile < p.n_tiles ? kGranPerTile : 0;
        }
        if (k > p.k_full) return 0;
        row0 = p.rem_row0 + 64ll * ((long long)b * p.rem_base + (b < p.rem_extra ? b : p.rem_extra));
        return p.rem_base + (b < p.rem_extra ? 1 : 0);
    };
    auto tile_at = [&](int k, long long &row0) -> int { return tile_of((int)blockIdx.x, k, row0); };
    // rounds of the cluster: the largest tile count among its CTAs
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kClusterMask = (uint16_t)((1u << CL) - 1u);
    int rounds = 0;
    for (int j = 0; j < CL; ++j) {
        int n = 0;
        long long r0;
        while (tile_of((int)blockIdx.x - (int)crank + j, n, r0) != 0) ++n;
        rounds = n > rounds ? n : rounds;
    }

    if (threadIdx.x == 0) {
        prefetch_tensormap(&tm_spec);
        prefetch_tensormap(&tm_spec64);
        prefetch_tensormap(&tm_w);
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), CL);  // one commit from the MMA warp of every CTA of the cluster
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tmem_relinquish();
    }
    for (int e = threadIdx.x; e < p.e_pad; e += kFitThreads) cs[e] = (e < p.n_comp) ? p.colscale[e] : 0.f;
    float *hs = cs + p.e_pad;  // G^-1 for the in-epilogue fix-up (only if p.hinv32)
    // staging area of the bulk stores: per epilogue warp [32 components][32 pixels] fp32, 128-byte aligned
    float *stage_all = reinterpret_cast<float *>(
        (reinterpret_cast<uintptr_t>(hs + (p.hinv32 ? p.n_comp * p.n_comp : 0)) + 127) & ~(uintptr_t)127);
    if (p.hinv32)
        for (int i = threadIdx.x; i < p.n_comp * p.n_comp; i += kFitThreads) hs[i] = p.hinv32[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (CL > 1) cluster_sync_all();  // every CTA's barriers exist before a partner multicasts into them

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < rounds; ++k) {
                long long row0 = 0;
                const int n_gran = tile_at(k, row0);  // 0: a round for the partners' sake (CL > 1 only)
                const uint32_t w_bytes = (uint32_t)(p.n_mma * kRowBytes);
                for (int kb = p.kb_begin; kb < p.kb_end; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t dst = ring + (uint32_t)stage * p.stage_bytes;
                    if (n_gran == kGranPerTile) {
                        mbar_arrive_expect_tx(full_bar(stage), p.stage_bytes - ((p.debug & 2) ? w_bytes : 0u));
                        tma_load_2d(dst, &tm_spec, full_bar(stage), kb * p.bke, (int)row0, kEvictFirst);
                    } else {
                        // partial tile: one 64-row box per granule (stacked boxes = the layout of a taller one: the
                        // 128-byte swizzle repeats every 8 rows)
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)n_gran * (kSubTileBytes / 2) + ((p.debug & 2) ? 0u : w_bytes));
                        for (int g = 0; g < n_gran; ++g)
                            tma_load_2d(dst + (uint32_t)g * (kSubTileBytes / 2), &tm_spec64, full_bar(stage), kb * p.bke,
                                        (int)row0 + 64 * g, kEvictFirst);
                    }
                    if (p.debug & 2) {
                        // (results are garbage: the MMAs read whatever the stage holds)
                    } else if (CL == 1) {
                        tma_load_2d(dst + M_TILES * kSubTileBytes, &tm_w, full_bar(stage), kb * p.bke, 0, kEvictLast);
                    } else {
                        const int part_rows = p.n_mma / CL;
                        tma_load_2d_multicast(dst + M_TILES * kSubTileBytes + crank * (uint32_t)(part_rows * kRowBytes), &tm_w,
                                              full_bar(stage), kb * p.bke, (int)crank * part_rows, kClusterMask, kEvictLast);
                    }
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int k = 0; k < rounds; ++k) {
                long long row0;
                const int n_gran = tile_at(k, row0);  // 0: nothing to multiply, the stages are only released
                const int n_sub = (n_gran + 1) >> 1;
                if (n_gran) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
                    tc_fence_after();
                }
                for (int kb = p.kb_begin; kb < p.kb_end; ++kb) {
                    mbar_wait(full_bar(stage), phase);  // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t a_addr = ring + (uint32_t)stage * p.stage_bytes;
                    const uint64_t w_desc = umma_desc_sw128(a_addr + M_TILES * kSubTileBytes);
#pragma unroll
                    for (int mt = 0; mt < M_TILES; ++mt) {
                        if (mt < n_sub && !(p.debug & 4)) {
                            const uint64_t a_desc = umma_desc_sw128(a_addr + mt * kSubTileBytes);
                            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols + (uint32_t)(mt * p.n_mma);
                            const uint32_t idesc = (mt == n_sub - 1 && (n_gran & 1)) ? p.idesc64 : p.idesc;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                // 32 bytes (16 fp16 / 8 tf32) along K per instruction: +2 in 16-byte units
                                tc_mma<KIND>(d_tmem, a_desc + (uint64_t)(2 * kk), w_desc + (uint64_t)(2 * kk), idesc,
                                             (kb > p.kb_begin || kk > 0) ? 1u : 0u);
                            }
                        }
                    }
                    // frees the smem stage once these MMAs retire -- in every ring of the cluster
                    if (CL == 1)
                        tc_commit(empty_bar(stage));
                    else
                        tc_commit_multicast(empty_bar(stage), kClusterMask);
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (n_gran) {
                    tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1u;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> x[e][p] =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int k = 0;; ++k) {
            long long row0;
            const int n_gran = tile_at(k, row0);
            if (n_gran == 0) break;
            const int n_sub = (n_gran + 1) >> 1;
            // pixel of this lane in sub-tile mt: a 128-row MMA fills all 32 lanes of the warp's TMEM quarter, the 64-row
            // MMA of an odd last granule only lanes 0..15 (rows 16 q .. 16 q + 15)
            auto pixel_of = [&](int mt, bool &valid) -> long long {
                const bool m64 = (mt == n_sub - 1) && (n_gran & 1);
                const long long px = row0 + mt * kTileRows + (m64 ? q * 16 + lane : q * 32 + lane);
                valid = mt < n_sub && (!m64 || lane < 16) && px < p.n_px;
                return px;
            };
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            if (p.reg_path) {
                // ---- whole pixels in registers: drain TMEM for both sub-tiles first and hand the
                // accumulator back to the MMA warp, then fix small zero sets and store ----
                float xv[M_TILES][32];
#pragma unroll
                for (int mt = 0; mt < M_TILES; ++mt) {
                    const uint32_t taddr =
                        tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(mt * p.n_mma);
#pragma unroll
                    for (int c0 = 0; c0 < 32; c0 += 8) {
                        if (c0 < p.e_pad && mt < n_sub) {
                            float a8[8], part[8];
                            tmem_ld8(taddr + (uint32_t)((p.n_planes - 1) * p.e_pad + c0), a8);
                            tmem_ld_wait();
                            for (int d = p.n_planes - 2; d >= 0; --d) {
                                tmem_ld8(taddr + (uint32_t)(d * p.e_pad + c0), part);
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 8; ++j) a8[j] = fmaf(a8[j], p.plane_ratio, part[j]);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) xv[mt][c0 + j] = cs[c0 + j] * a8[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) xv[mt][c0 + j] = 0.f;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));
#pragma unroll
                for (int mt = 0; mt < M_TILES; ++mt) {
                    bool valid;
                    const long long px = pixel_of(mt, valid);
                    bool any_neg = false;
#pragma unroll
                    for (int e = 0; e < 32; ++e) any_neg |= (xv[mt][e] < 0.f);
                    bool negative = any_neg && valid;
                    if (negative && p.hinv32 != nullptr && !(p.debug & 8)) negative = !fixup_small(xv[mt], p.n_comp, hs);
                    // Whole groups of 32 pixels leave through shared memory: every lane parks its pixel's values in the
                    // warp's staging tile and lane e sends component e's 128 bytes with one cp.async.bulk -- the warp does
                    // not wait for 32 global stores per sub-tile (which cost the epilogue 29 us per tile and kept it
                    // level with the MMAs) but for nothing.
                    const bool m64 = (mt == n_sub - 1) && (n_gran & 1);
                    const bool bulk = p.bulk && !m64 && !(p.debug & 1) && __all_sync(0xffffffffu, valid);
                    if (bulk) {
                        float *st = stage_all + (warp - 2) * 1024;
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // my earlier row has been read
                        __syncwarp();
#pragma unroll
                        for (int e = 0; e < 32; ++e) st[e * 32 + lane] = xv[mt][e];
                        fence_proxy_async();
                        __syncwarp();
                        if (lane < p.n_comp) {
                            // component-owner exchange: the row goes to the GPU that assembles the map of component e
                            // (row pointer already offset to this rank's pixels; NULL: nobody wants it).  Values of pixels
                            // that still go to the general solver travel too; that kernel, next in the stream, overwrites
                            // them with the final ones.
                            float *dst = p.peers.n_owned > 0 ? p.peers.owner_row[lane]
                                                             : p.x + (long long)lane * p.ld_x;
                            if (dst != nullptr) {
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(dst + (px - lane)),
                                             "r"(smem_u32(st + lane * 32))
                                             : "memory");
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                        // owner mode: the local array carries the pixels the general solver still has to read
                        if (p.peers.n_owned > 0 && negative) {
#pragma unroll
                            for (int e = 0; e < 32; ++e)
                                if (e < p.n_comp) p.x[(long long)e * p.ld_x + px] = xv[mt][e];
                        }
                    }
                    if (valid && !bulk && !(p.debug & 1)) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            if (e < p.n_comp) {
                                const long long at = (long long)e * p.ld_x + px;
                                // component-owner exchange: final values live on the owners only; the local array
                                // carries just the pixels that still go to the general solver (it reads them there).
                                // One store per value instead of two: the epilogue has to keep pace with the MMAs.
                                if (p.peers.n_owned == 0 || negative) p.x[at] = xv[mt][e];
                                // Remote copies only of FINAL values: a pixel that still goes to the general
                                // solver is pushed by that kernel, so no destination ever sees two stores
                                // to the same address from different kernels.
                                if (!negative || p.neg_list == nullptr) {  // (no list: no solver follows)
                                    if (p.peers.n_owned > 0) {
                                        // component-owner exchange: the value goes to the GPU that assembles the
                                        // map of component e (row pointer already offset to this rank's pixels)
                                        float *row = p.peers.owner_row[e];
                                        if (row) row[px] = xv[mt][e];
                                    }
                                    for (int k = 0; k < p.peers.n_peers; ++k) p.peers.peer[k][at] = xv[mt][e];
                                    if (p.peers.multicast)
                                        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.peers.multicast + at),
                                                     "f"(xv[mt][e])
                                                     : "memory");
                                }
                            }
                        }
                    }
                    const unsigned ballot = __ballot_sync(0xffffffffu, negative);
                    if (ballot != 0u && p.neg_list != nullptr) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(p.neg_count, __popc(ballot));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (negative) p.neg_list[base + __popc(ballot & ((1u << lane) - 1u))] = (int)(p.px_offset + px);
                    }
                }
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
                continue;
            }
#pragma unroll
            for (int mt = 0; mt < M_TILES; ++mt) {
                if (mt >= n_sub) continue;
                bool valid;
                const long long px = pixel_of(mt, valid);
                const uint32_t taddr =
                    tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_cols + (uint32_t)(mt * p.n_mma);
                bool negative = false;
                for (int c0 = 0; c0 < p.e_pad; c0 += 8) {
                    // planes are stacked along N; Horner from the least significant plane:
                    // x = cs * (S0 + r * (S1 + r * S2))
                    float acc[8], part[8];
                    tmem_ld8(taddr + (uint32_t)((p.n_planes - 1) * p.e_pad + c0), acc);
                    tmem_ld_wait();
                    for (int d = p.n_planes - 2; d >= 0; --d) {
                        tmem_ld8(taddr + (uint32_t)(d * p.e_pad + c0), part);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[j] = fmaf(acc[j], p.plane_ratio, part[j]);
                    }
                    if (valid) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = cs[c0 + j] * acc[j];  // cs is 0 for padding columns
                            negative |= (v[j] < 0.f);
                        }
                        if (p.pixel_major) {
                            // one record per pixel: two 128-bit stores per 8 components, also to the peers
                            const long long at = px * p.ld_x + c0;
                            const float4 a = make_float4(v[0], v[1], v[2], v[3]), b = make_float4(v[4], v[5], v[6], v[7]);
                            *reinterpret_cast<float4 *>(p.x + at) = a;
                            *reinterpret_cast<float4 *>(p.x + at + 4) = b;
                            for (int k = 0; k < p.peers.n_peers; ++k) {
                                *reinterpret_cast<float4 *>(p.peers.peer[k] + at) = a;
                                *reinterpret_cast<float4 *>(p.peers.peer[k] + at + 4) = b;
                            }
                            if (p.peers.multicast) {
                                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                                 p.peers.multicast + at),
                                             "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w)
                                             : "memory");
                                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                                 p.peers.multicast + at + 4),
                                             "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                                             : "memory");
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int e = c0 + j;
                                if (e < p.n_comp) {
                                    const long long at = (long long)e * p.ld_x + px;
                                    p.x[at] = v[j];
                                    // fused gather of the maps: same value to the peers' buffers over NVLink
                                    for (int k = 0; k < p.peers.n_peers; ++k) p.peers.peer[k][at] = v[j];
                                    if (p.peers.multicast)
                                        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.peers.multicast + at),
                                                     "f"(v[j])
                                                     : "memory");
                                }
                            }
                        }
                    }
                }
                if (p.neg_list != nullptr) {
                    // compact the pixels that need the non-negative fix-up: one atomic per warp
                    const unsigned ballot = __ballot_sync(0xffffffffu, negative);
                    if (ballot != 0u) {
                        int base = 0;
                        if (lane == 0) base = atomicAdd(p.neg_count, __popc(ballot));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (negative) p.neg_list[base + __popc(ballot & ((1u << lane) - 1u))] = (int)(p.px_offset + px);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    if (p.bulk && warp >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();  // no partner may still write into this CTA's stages or barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---- CUDA-core version of the same contraction (cross-check / TMA-unfriendly layouts) ---------
template <typename T>
__device__ __forceinline__ float ld_val(const T *p);
template <>
__device__ __forceinline__ float ld_val<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_val<__half>(const __half *p) { return __half2float(__ldg(p)); }

constexpr int kSimtPx = 4;    // pixels per warp
constexpr int kSimtComp = 16; // components per pass

template <typename T>
__global__ void __launch_bounds__(256)
fit_contract_simt_kernel(const T *__restrict__ spec, long long n_px, long long ld_spec, int k_begin, int k_end,
                         const float *__restrict__ w, long long ld_w, int n_comp, float *__restrict__ x,
                         long long ld_x, int pixel_major) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long px0 = warp_global * kSimtPx;
    if (px0 >= n_px) return;
    for (int e0 = 0; e0 < n_comp; e0 += kSimtComp) {
        float acc[kSimtPx][kSimtComp];
#pragma unroll
        for (int i = 0; i < kSimtPx; ++i)
#pragma unroll
            for (int j = 0; j < kSimtComp; ++j) acc[i][j] = 0.f;
        for (int k = k_begin + lane; k < k_end; k += 32) {
            float y[kSimtPx];
#pragma unroll
            for (int i = 0; i < kSimtPx; ++i)
                y[i] = (px0 + i < n_px) ? ld_val<T>(spec + (px0 + i) * ld_spec + k) : 0.f;
#pragma unroll
            for (int j = 0; j < kSimtComp; ++j) {
                const float wv = (e0 + j < n_comp) ? __ldg(w + (long long)(e0 + j) * ld_w + k) : 0.f;
#pragma unroll
                for (int i = 0; i < kSimtPx; ++i) acc[i][j] = fmaf(y[i], wv, acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < kSimtPx; ++i)
#pragma unroll
            for (int j = 0; j < kSimtComp; ++j) {
                float v = acc[i][j];
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
                if (lane == 0 && px0 + i < n_px && e0 + j < n_comp)
                    x[pixel_major ? (px0 + i) * ld_x + (e0 + j) : (long long)(e0 + j) * ld_x + px0 + i] = v;
            }
    }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// L2 promotion of the TMA loads: 256 B by default; MT_FIT_L2PROMO=0|1|2|3 selects none / 64 / 128 / 256 B (tuning knob)
CUtensorMapL2promotion l2_promotion() {
    const char *env = getenv("MT_FIT_L2PROMO");
    switch (env ? atoi(env) : 3) {
        case 0: return CU_TENSOR_MAP_L2_PROMOTION_NONE;
        case 1: return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
        case 2: return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        default: return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }
}

// 2-D row-major tensor [rows][cols] with a row pitch, boxes of [box_rows][128 bytes], swizzle 128B
int make_tile_map(CUtensorMap *map, const void *base, int dtype, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                  uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return mt::fail(MT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const uint32_t es = dtype == MT_DTYPE_F16 ? 2u : 4u;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_elems * es};
    cuuint32_t box[2] = {kRowBytes / es, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype == MT_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return mt::fail(MT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return MT_OK;
}

template <int KIND, int M_TILES, int CL>
int launch_fit(const CUtensorMap &tm_spec, const CUtensorMap &tm_spec64, const CUtensorMap &tm_w, const FitParams &p,
               size_t smem, int grid, cudaStream_t stream) {
    auto kern = fit_contract_kernel<KIND, M_TILES, CL>;
    static int smem_set[16] = {0};  // per device: opt in to the large dynamic smem carve-out once
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
        if (dev < 16) smem_set[dev] = kSmemLimit;
    }
    if (CL == 1) {
        kern<<<grid, kFitThreads, smem, stream>>>(tm_spec, tm_spec64, tm_w, p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kFitThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm_spec, tm_spec64, tm_w, p));
    }
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}

// CTAs of clusters of CL that can be resident at once with this kernel's shared memory (one CTA per SM): the
// GPCs do not all hold a multiple of CL SMs.  0 if the query fails.
template <int KIND, int M_TILES, int CL>
int resident_cluster_ctas(size_t smem) {
    auto kern = fit_contract_kernel<KIND, M_TILES, CL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(mt::num_sms() / CL * CL));
    cfg.blockDim = dim3(kFitThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n * CL;
}

}  // namespace

extern "C" int mt_fit_contract(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec,
                               int32_t n_ch_total, int32_t k_begin, int32_t k_end, const void *wpack_dev,
                               int64_t ld_w, int32_t e_pad, int32_t n_comp, int32_t n_planes,
                               const float *colscale_dev, float plane_ratio, float *x_dev, int64_t ld_x,
                               int32_t x_layout, int64_t px_offset, int32_t *neg_list_dev, int32_t *neg_count_dev,
                               const float *hinv32_dev, const MtPeers *peers_host, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && wpack_dev && colscale_dev && x_dev, "mt_fit_contract: null pointer");
    MT_REQUIRE(dtype == MT_DTYPE_F32 || dtype == MT_DTYPE_F16, "mt_fit_contract: dtype %d", dtype);
    const int es = dtype == MT_DTYPE_F16 ? 2 : 4;
    MT_REQUIRE(n_px >= 0 && n_px < (1ll << 31) - 512, "mt_fit_contract: n_px=%lld", (long long)n_px);
    MT_REQUIRE(n_ch_total > 0 && ld_spec >= n_ch_total && ld_w >= n_ch_total,
               "mt_fit_contract: n_ch_total=%d ld_spec=%lld ld_w=%lld", n_ch_total, (long long)ld_spec,
               (long long)ld_w);
    MT_REQUIRE(0 <= k_begin && k_begin < k_end && k_end <= n_ch_total, "mt_fit_contract: channel window [%d,%d)",
               k_begin, k_end);
    MT_REQUIRE(n_planes >= 1 && n_planes <= 4, "mt_fit_contract: n_planes=%d", n_planes);
    MT_REQUIRE(e_pad >= 8 && e_pad % 8 == 0 && n_comp >= 1 && n_comp <= e_pad && (n_planes * e_pad) % 16 == 0 &&
                   n_planes * e_pad <= 256,
               "mt_fit_contract: e_pad=%d n_comp=%d n_planes=%d (need n_planes*e_pad a multiple of 16, <= 256)",
               e_pad, n_comp, n_planes);
    if (x_layout == MT_X_PIXEL_MAJOR)
        MT_REQUIRE(ld_x >= e_pad && ld_x % 4 == 0 && ((uintptr_t)x_dev & 15) == 0,
                   "mt_fit_contract: pixel-major x needs ld_x >= e_pad, ld_x %% 4 == 0 and a 16-byte aligned base");
    else
        MT_REQUIRE(x_layout == MT_X_ELEMENT_MAJOR && ld_x >= n_px, "mt_fit_contract: ld_x=%lld < n_px", (long long)ld_x);
    if ((ld_spec * es) % 16 || (ld_w * es) % 16 || ((uintptr_t)spec_dev & 15) || ((uintptr_t)wpack_dev & 15))
        return mt::fail(MT_ERR_UNSUPPORTED,
                        "mt_fit_contract: TMA needs 16-byte aligned base pointers and row pitches "
                        "(ld_spec=%lld ld_w=%lld elements of %d bytes); use mt_fit_contract_simt",
                        (long long)ld_spec, (long long)ld_w, es);
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);

    FitParams p{};
    MT_REQUIRE((neg_list_dev == nullptr) == (neg_count_dev == nullptr),
               "mt_fit_contract: neg_list and neg_count must be given together");
    MT_REQUIRE(px_offset >= 0 && px_offset + n_px < (1ll << 31), "mt_fit_contract: px_offset=%lld", (long long)px_offset);
    p.n_px = n_px;
    p.ld_x = ld_x;
    p.px_offset = px_offset;
    p.pixel_major = x_layout == MT_X_PIXEL_MAJOR;
    if (hinv32_dev) {
        MT_REQUIRE(e_pad <= 32 && x_layout == MT_X_ELEMENT_MAJOR && neg_list_dev,
                   "mt_fit_contract: the in-epilogue fix-up needs e_pad <= 32, element-major x and a fallback list");
        p.hinv32 = hinv32_dev;
    }
    p.reg_path = (hinv32_dev != nullptr) || (peers_host && peers_host->n_owned > 0);
    p.neg_list = neg_list_dev;
    p.neg_count = neg_count_dev;
    if (peers_host) {
        MT_REQUIRE(peers_host->n_peers >= 0 && peers_host->n_peers <= MT_MAX_PEERS, "mt_fit_contract: n_peers=%d",
                   peers_host->n_peers);
        MT_REQUIRE(peers_host->n_owned == 0 || (e_pad <= 32 && x_layout == MT_X_ELEMENT_MAJOR),
                   "mt_fit_contract: per-component destinations need n_comp <= 32 and element-major x");
        p.peers = *peers_host;
    }
    p.x = x_dev;
    p.colscale = colscale_dev;
    p.bke = kRowBytes / es;
    p.kb_begin = k_begin / p.bke;
    p.kb_end = (k_end + p.bke - 1) / p.bke;
    p.n_mma = n_planes * e_pad;
    p.e_pad = e_pad;
    p.n_comp = n_comp;
    p.n_planes = n_planes;
    p.plane_ratio = plane_ratio;
    const int m_tiles = (p.n_mma <= 128) ? 2 : 1;
    const int block_rows = kTileRows * m_tiles;
    p.n_tiles = (int)((n_px + block_rows - 1) / block_rows);
    p.stage_bytes = (uint32_t)(m_tiles * kSubTileBytes + p.n_mma * kRowBytes);
    // bulk stores of the epilogue (MT_FIT_BULK_STORE=0 restores scalar stores): register path without peer destinations,
    // 16-byte aligned component rows
    static const bool bulk_pref = [] {
        const char *v = getenv("MT_FIT_BULK_STORE");
        return !(v && atoi(v) == 0);
    }();
    p.bulk = bulk_pref && p.reg_path && p.peers.n_peers == 0 && p.peers.multicast == nullptr &&
             x_layout == MT_X_ELEMENT_MAJOR && (ld_x % 4) == 0 && ((uintptr_t)x_dev & 15) == 0;
    for (int e = 0; e < n_comp && e < MT_MAX_OWNED && p.bulk && p.peers.n_owned > 0; ++e)
        p.bulk = ((uintptr_t)p.peers.owner_row[e] & 15) == 0;  // (NULL rows pass)
    const size_t fixed = 1024 /*alignment slack*/ + 8 * (2 * 12 + 4) + 16 + 4 * (size_t)e_pad + 64 +
                         (hinv32_dev ? 4 * (size_t)n_comp * n_comp : 0) + (p.bulk ? 128 + 4 * 4096 : 0);
    p.n_stages = (int)((kSmemLimit - fixed) / p.stage_bytes);
    if (p.n_stages > 12) p.n_stages = 12;
    if (const char *cap = getenv("MT_FIT_MAX_STAGES")) {  // tuning knob (e.g. to leave shared memory to a co-resident kernel)
        const int c = atoi(cap);
        if (c >= 2 && c < p.n_stages) p.n_stages = c;
    }
    MT_REQUIRE(p.n_stages >= 2, "mt_fit_contract: tile does not fit shared memory");
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * m_tiles * p.n_mma)) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = umma_idesc(dtype == MT_DTYPE_F16 ? 0 : 2, kTileRows, p.n_mma);
    const size_t smem = (size_t)p.n_stages * p.stage_bytes + fixed;

    const int sms = mt::num_sms();
    // Tile schedule (see the kernel): whole tiles round-robin, then the remainder in 64-pixel granules, one partial tile
    // per CTA.  MT_FIT_TAIL=0 restores whole tiles only (the last wave then runs partly empty).
    static const bool tail = [] {
        const char *v = getenv("MT_FIT_TAIL");
        return !(v && atoi(v) == 0);
    }();
    // Clusters of 2 CTAs share the pack slabs by TMA multicast (MT_FIT_CLUSTER=1 switches it off): only for problems
    // that give every CTA work, pack halves of whole 8-row swizzle atoms, and where (nearly) all SMs can host a cluster
    static const int cluster_pref = [] {
        const char *v = getenv("MT_FIT_CLUSTER");
        return v ? atoi(v) : 2;
    }();
    const long long n_gran = (n_px + 63) / 64;
    const int gran_per_tile = 2 * m_tiles;
    int grid = sms;
    int cl = 1;
    if (tail && cluster_pref == 2 && n_gran >= 2ll * gran_per_tile * sms && (p.n_mma / 2) % 8 == 0) {
        static int cached[2][2][16];  // [dtype][m_tiles - 1][device]: resident cluster CTAs + 1 (0 = not asked yet)
        int dev = 0;
        cudaGetDevice(&dev);
        int &slot = cached[dtype == MT_DTYPE_F16 ? 0 : 1][m_tiles - 1][dev & 15];
        if (slot == 0) {
            const size_t smem_guess = (size_t)kSmemLimit;
            int n = 0;
            if (dtype == MT_DTYPE_F16)
                n = m_tiles == 2 ? resident_cluster_ctas<0, 2, 2>(smem_guess) : resident_cluster_ctas<0, 1, 2>(smem_guess);
            else
                n = m_tiles == 2 ? resident_cluster_ctas<1, 2, 2>(smem_guess) : resident_cluster_ctas<1, 1, 2>(smem_guess);
            slot = n + 1;
        }
        const int resident = slot - 1;
        if (resident >= sms - 4 && resident >= 2) {
            cl = 2;
            grid = resident < sms ? resident : (sms / 2) * 2;
        }
    }
    if (const char *cap = getenv("MT_FIT_GRID")) {  // measurement knob: fewer CTAs than SMs
        const int c = atoi(cap);
        if (c >= 2 && c < grid) grid = cl == 2 ? (c / 2) * 2 : c;
    }
    if (tail) {
        if (cl == 1) grid = (int)(n_gran < grid ? n_gran : grid);
        p.k_full = (int)(n_gran / ((long long)gran_per_tile * grid));
        const long long rem = n_gran - (long long)p.k_full * gran_per_tile * grid;  // < gran_per_tile * grid
        p.rem_base = (int)(rem / grid);
        p.rem_extra = (int)(rem % grid);
        p.rem_row0 = (long long)p.k_full * grid * block_rows;
    } else {
        grid = p.n_tiles < grid ? p.n_tiles : grid;
        p.k_full = (p.n_tiles + grid - 1) / grid;
        p.rem_base = p.rem_extra = 0;
        p.rem_row0 = 0;
    }
    p.idesc64 = umma_idesc(dtype == MT_DTYPE_F16 ? 0 : 2, 64, p.n_mma);
    {
        static const int dbg = [] {
            const char *v = getenv("MT_FIT_DEBUG");
            return v ? atoi(v) : 0;
        }();
        p.debug = dbg;
    }

    CUtensorMap tm_spec, tm_spec64, tm_w;
    if (int rc = make_tile_map(&tm_spec, spec_dev, dtype, (uint64_t)n_ch_total, (uint64_t)n_px, (uint64_t)ld_spec,
                               (uint32_t)block_rows))
        return rc;
    if (int rc = make_tile_map(&tm_spec64, spec_dev, dtype, (uint64_t)n_ch_total, (uint64_t)n_px, (uint64_t)ld_spec, 64u))
        return rc;
    if (int rc = make_tile_map(&tm_w, wpack_dev, dtype, (uint64_t)n_ch_total, (uint64_t)p.n_mma, (uint64_t)ld_w,
                               (uint32_t)(p.n_mma / cl)))
        return rc;
#define MT_FIT_LAUNCH(K, M)                                                                              \
    (cl == 2 ? launch_fit<K, M, 2>(tm_spec, tm_spec64, tm_w, p, smem, grid, stream)                         \
             : launch_fit<K, M, 1>(tm_spec, tm_spec64, tm_w, p, smem, grid, stream))
    int rc_launch;
    if (dtype == MT_DTYPE_F16)
        rc_launch = m_tiles == 2 ? MT_FIT_LAUNCH(0, 2) : MT_FIT_LAUNCH(0, 1);
    else
        rc_launch = m_tiles == 2 ? MT_FIT_LAUNCH(1, 2) : MT_FIT_LAUNCH(1, 1);
    return rc_launch;
#undef MT_FIT_LAUNCH
}

extern "C" int mt_fit_contract_simt(const void *spec_dev, int32_t dtype, int64_t n_px, int64_t ld_spec,
                                    int32_t k_begin, int32_t k_end, const float *w_dev, int64_t ld_w,
                                    int32_t n_comp, float *x_dev, int64_t ld_x, int32_t x_layout, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(spec_dev && w_dev && x_dev, "mt_fit_contract_simt: null pointer");
    MT_REQUIRE(n_px >= 0 && 0 <= k_begin && k_begin < k_end && k_end <= ld_spec && k_end <= ld_w && n_comp >= 1 &&
                   ld_x >= (x_layout == MT_X_PIXEL_MAJOR ? (int64_t)n_comp : n_px),
               "mt_fit_contract_simt: bad sizes");
    const int pm = x_layout == MT_X_PIXEL_MAJOR;
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long warps = (n_px + kSimtPx - 1) / kSimtPx;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    if (dtype == MT_DTYPE_F32)
        fit_contract_simt_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float *>(spec_dev), n_px, ld_spec,
                                                                  k_begin, k_end, w_dev, ld_w, n_comp, x_dev, ld_x, pm);
    else if (dtype == MT_DTYPE_F16)
        fit_contract_simt_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half *>(spec_dev), n_px,
                                                                   ld_spec, k_begin, k_end, w_dev, ld_w, n_comp,
                                                                   x_dev, ld_x, pm);
    else
        return mt::fail(MT_ERR_UNSUPPORTED, "mt_fit_contract_simt: dtype %d", dtype);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
