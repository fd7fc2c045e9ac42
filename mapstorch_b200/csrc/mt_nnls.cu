// Batched per-pixel non-negative solve on the shared normal equations.
//
// The reference parametrises amplitudes as 10**a (opt.py:154, map.py:163), i.e. it solves a
// non-negative least-squares problem; with fixed global parameters its converged answer is
//     argmin_x (x - xu)^T G (x - xu)   s.t. x >= 0,      G = B^T B  (pixel independent),
// where xu is the unconstrained solution produced by mt_fit_contract.  One thread owns one
// pixel and runs block principal pivoting in fp64; each pivot step solves an SPD system on the
// smaller of the two index sets:
//     zero set  A small:  z = (H_AA)^-1 xu_A,  x_F = xu_F - H_FA z,  dual y_A = -z      (H = G^-1)
//     free set  F small:  x_F = (G_FF)^-1 b_F,                       dual y_A = G_AF x_F - b_A
// so the Cholesky factor never exceeds (E/2)(E/2+1)/2 entries of thread-local memory.  Pixels
// whose xu is already non-negative (the common case on high-count maps) only pay the scan.
#include "mt_common.cuh"

namespace {

constexpr int kNnlsThreads = 128;

// pixels whose pivot system turned out numerically singular (near-collinear components): they receive the clamped
// unconstrained solution and are counted here; the host reads the total with mt_nnls_singular_count()
__device__ unsigned long long g_singular_pixels = 0ull;

// Packed lower-triangular Cholesky of M[idx, idx] (M: shared, row-major ld), then solve into v.
// Returns false if a pivot is not positive (numerically singular subset).
template <int HALF>
__device__ bool chol_solve(const double *M, int ld, const unsigned char *idx, int n, double *L, double *v) {
    for (int j = 0; j < n; ++j) {
        const double *mj = M + idx[j] * ld;
        for (int i = 0; i <= j; ++i) {
            double s = mj[idx[i]];
            const double *lj = L + j * (j + 1) / 2, *li = L + i * (i + 1) / 2;
            for (int k = 0; k < i; ++k) s -= lj[k] * li[k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[j * (j + 1) / 2 + j] = sqrt(s);
            } else {
                L[j * (j + 1) / 2 + i] = s / li[i];
            }
        }
    }
    for (int i = 0; i < n; ++i) {  // forward
        double s = v[i];
        const double *li = L + i * (i + 1) / 2;
        for (int k = 0; k < i; ++k) s -= li[k] * v[k];
        v[i] = s / li[i];
    }
    for (int i = n - 1; i >= 0; --i) {  // backward
        double s = v[i];
        for (int k = i + 1; k < n; ++k) s -= L[k * (k + 1) / 2 + i] * v[k];
        v[i] = s / L[i * (i + 1) / 2 + i];
    }
    return true;
}

template <int EMAX>
__device__ void fix_pixel(float *__restrict__ x, long long ld_x, int pixel_major, long long px, int E, const double *G,
                          const double *H, int *__restrict__ n_fixed, const MtPeers &peers);

template <int EMAX>
__global__ void __launch_bounds__(kNnlsThreads)
nnls_fixup_kernel(float *__restrict__ x, long long ld_x, int pixel_major, long long n_px, int E,
                  const double *__restrict__ g_dev,
                  const double *__restrict__ h_dev, int *__restrict__ n_fixed, const int *__restrict__ list,
                  const int *__restrict__ list_count, int list_min, const MtPeers peers) {
    extern __shared__ double mats[];  // G[E*E], H[E*E]
    // list mode: blocks beyond the (device-side) item count have nothing to do; lists of at most list_min items
    // belong to the warp-per-pixel kernel
    if (list && ((long long)blockIdx.x * kNnlsThreads >= (long long)*list_count || *list_count <= list_min)) return;
    double *G = mats, *H = mats + E * E;
    for (int i = threadIdx.x; i < E * E; i += kNnlsThreads) {
        G[i] = g_dev[i];
        H[i] = h_dev[i];
    }
    __syncthreads();
    // scan mode: one thread per pixel of [0, n_px); list mode: grid-stride over the compacted list
    const long long n_items = list ? (long long)*list_count : n_px;
    for (long long item = (long long)blockIdx.x * kNnlsThreads + threadIdx.x; item < n_items;
         item += (long long)gridDim.x * kNnlsThreads)
        fix_pixel<EMAX>(x, ld_x, pixel_major, list ? (long long)list[item] : item, E, G, H, n_fixed, peers);
}

template <int EMAX>
__device__ void fix_pixel(float *__restrict__ x, long long ld_x, int pixel_major, long long px, int E, const double *G,
                          const double *H, int *__restrict__ n_fixed, const MtPeers &peers) {
    const long long se = pixel_major ? 1 : ld_x, base = pixel_major ? px * ld_x : px;  // x[base + e * se]
    constexpr int HALF = EMAX / 2;
    double xu[EMAX];
    bool any_neg = false;
    double xmax = 0.0;
    for (int e = 0; e < E; ++e) {
        xu[e] = (double)x[base + e * se];
        any_neg |= (xu[e] < 0.0);
        xmax = fmax(xmax, fabs(xu[e]));
    }
    if (!any_neg) return;
    if (n_fixed) atomicAdd(n_fixed, 1);

    double xs[EMAX];        // current primal (free set) / dual (zero set) values
    double b[EMAX];         // G xu, filled lazily
    bool have_b = false;
    bool free_[EMAX];
    unsigned char idx[EMAX];
    double L[HALF * (HALF + 1) / 2];
    double v[HALF];
    int n_free = 0;
    for (int e = 0; e < E; ++e) {
        free_[e] = xu[e] > 0.0;
        n_free += free_[e];
        xs[e] = xu[e];
    }
    double gmax = 0.0;
    for (int e = 0; e < E; ++e) gmax = fmax(gmax, G[e * E + e]);
    const double tol_x = 1e-12 * xmax, tol_y = 1e-12 * xmax * gmax;

    int ninf = E + 1, backup = 3;
    const int max_iter = 8 * E + 16;
    for (int iter = 0; iter < max_iter; ++iter) {
        const int n_zero = E - n_free;
        bool ok = true;
        if (n_zero <= n_free) {
            // zero set is the smaller one: Schur complement on H = G^-1
            int n = 0;
            for (int e = 0; e < E; ++e)
                if (!free_[e]) {
                    idx[n] = (unsigned char)e;
                    v[n] = xu[e];
                    ++n;
                }
            ok = chol_solve<HALF>(H, E, idx, n, L, v);
            if (ok) {
                for (int e = 0; e < E; ++e) {
                    if (free_[e]) {
                        double s = xu[e];
                        const double *he = H + e * E;
                        for (int k = 0; k < n; ++k) s -= he[idx[k]] * v[k];
                        xs[e] = s;
                    }
                }
                for (int k = 0; k < n; ++k) xs[idx[k]] = -v[k];  // dual on the zero set; tol_y scale: G*x
                // H-scaled dual has units x / H; compare on the same footing as G-form below
            }
        } else {
            if (!have_b) {
                for (int e = 0; e < E; ++e) {
                    double s = 0.0;
                    const double *ge = G + e * E;
                    for (int k = 0; k < E; ++k) s += ge[k] * xu[k];
                    b[e] = s;
                }
                have_b = true;
            }
            int n = 0;
            for (int e = 0; e < E; ++e)
                if (free_[e]) {
                    idx[n] = (unsigned char)e;
                    v[n] = b[e];
                    ++n;
                }
            ok = chol_solve<HALF>(G, E, idx, n, L, v);
            if (ok) {
                for (int k = 0; k < n; ++k) xs[idx[k]] = v[k];
                for (int e = 0; e < E; ++e) {
                    if (!free_[e]) {
                        double s = -b[e];
                        const double *ge = G + e * E;
                        for (int k = 0; k < n; ++k) s += ge[idx[k]] * v[k];
                        xs[e] = s;
                    }
                }
            }
        }
        if (!ok) {
            // singular subset (strongly overlapping components): fall back to the clamped unconstrained solution
            // rather than pairing stale values with a half-updated index set, and tell the host
            for (int e = 0; e < E; ++e) {
                free_[e] = xu[e] > 0.0;
                xs[e] = xu[e];
            }
            atomicAdd(&g_singular_pixels, 1ull);
            break;
        }
        int n_bad = 0, last_bad = -1;
        for (int e = 0; e < E; ++e) {
            const bool bad = free_[e] ? (xs[e] < -tol_x) : (xs[e] < -tol_y);
            if (bad) {
                ++n_bad;
                last_bad = e;
            }
        }
        if (n_bad == 0) break;
        bool flip_all;
        if (n_bad < ninf) {
            ninf = n_bad;
            backup = 3;
            flip_all = true;
        } else if (backup > 0) {
            --backup;
            flip_all = true;
        } else {
            flip_all = false;
        }
        for (int e = 0; e < E; ++e) {
            const bool bad = free_[e] ? (xs[e] < -tol_x) : (xs[e] < -tol_y);
            if (bad && (flip_all || e == last_bad)) {
                n_free += free_[e] ? -1 : 1;
                free_[e] = !free_[e];
            }
        }
    }
    if (pixel_major) {
        // one 16-byte aligned record per pixel: 128-bit stores, also to the peers / multicast mapping
        for (int e0 = 0; e0 < E; e0 += 4) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = (e0 + j < E && free_[e0 + j]) ? (float)fmax(xs[e0 + j], 0.0) : 0.f;
            const float4 v = make_float4(o[0], o[1], o[2], o[3]);
            const long long at = base + e0;
            *reinterpret_cast<float4 *>(x + at) = v;
            for (int k = 0; k < peers.n_peers; ++k) *reinterpret_cast<float4 *>(peers.peer[k] + at) = v;
            if (peers.multicast)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(peers.multicast + at),
                             "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                             : "memory");
        }
        return;
    }
    for (int e = 0; e < E; ++e) {
        const float out = (float)(free_[e] ? fmax(xs[e], 0.0) : 0.0);
        const long long at = base + e * se;
        x[at] = out;
        if (peers.n_owned > 0 && e < MT_MAX_OWNED && peers.owner_row[e]) peers.owner_row[e][px] = out;
        for (int k = 0; k < peers.n_peers; ++k) peers.peer[k][at] = out;  // keep the peers' copies in step
        if (peers.multicast)
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(peers.multicast + at), "f"(out) : "memory");
    }
}

template <int EMAX>
int launch_nnls(float *x, int64_t ld_x, int pixel_major, int64_t n_px, int E, const double *g, const double *h, int *n_fixed,
                const int *list, const int *list_count, int list_min, const MtPeers &peers, cudaStream_t stream) {
    auto kern = nnls_fixup_kernel<EMAX>;
    const size_t smem = sizeof(double) * 2 * (size_t)E * E;
    static int smem_set[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 16) smem_set[dev] = (int)smem;
    }
    const unsigned grid = list ? (unsigned)(mt::num_sms() * 8) : (unsigned)((n_px + kNnlsThreads - 1) / kNnlsThreads);
    kern<<<grid, kNnlsThreads, smem, stream>>>(x, ld_x, pixel_major, n_px, E, g, h, n_fixed, list, list_count, list_min,
                                               peers);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}


// ---- warp-per-pixel variant for SHORT lists ---------------------------------------------------
// When the K3 epilogue has already fixed the pixels with small zero sets, only a few hundred
// pixels per GPU are left and the latency of one solve is what counts: here one warp owns a pixel,
// lane e owns component e (E <= 32), the zero set's system H_AA z = xu_A is solved by Gauss-Jordan
// with row r in lane r's registers, x_F = xu_F - H_FA z, duals on A are -z.  fp64, registers and
// shuffles only; H = G^-1 in shared memory.
constexpr int kListWarps = 4;
// Lists longer than this go to the thread-per-pixel kernel instead (throughput, not latency, decides then: on the
// SNIP workload of BASELINE config 3 most pixels clamp three or more components and the warp kernel took 0.79 ms
// per 131 072-pixel shard).  The count lives on the device, so both kernels are launched and one returns at once.
constexpr int kShortListMax = 6144;

// Row `lane` of the Gauss-Jordan solve of H[a_r, a_c] z = v over the n <= NMAX zero-set components (a = this lane's
// component, v = its right-hand side): returns z of this row (0 for lanes >= n).  Registers and shuffles only.
template <int NMAX>
__device__ __forceinline__ double gauss_jordan_row(const double *Hs, int E, int a, int n, int lane, double v) {
    double M[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        M[c] = 0.0;
        if (c < n) {
            const int ac = __shfl_sync(0xffffffffu, a, c);
            M[c] = (lane < n) ? Hs[a * E + ac] : 0.0;
        }
    }
#pragma unroll
    for (int k = 0; k < NMAX; ++k) {
        if (k < n) {
            const double pkk = __shfl_sync(0xffffffffu, M[k], k);
            const double pv = __shfl_sync(0xffffffffu, v, k);
            const double f = (lane != k && lane < n) ? M[k] / pkk : 0.0;
#pragma unroll
            for (int c = k + 1; c < NMAX; ++c) {
                if (c < n) {
                    const double pc = __shfl_sync(0xffffffffu, M[c], k);
                    M[c] -= f * pc;
                }
            }
            v -= f * pv;
        }
    }
    double diag = 1.0;
#pragma unroll
    for (int c = 0; c < NMAX; ++c)
        if (c == lane) diag = M[c];
    return (lane < n) ? v / diag : 0.0;
}

__global__ void __launch_bounds__(kListWarps * 32)
nnls_warp_kernel(float *__restrict__ x, long long ld_x, const int *__restrict__ list,
                 const int *__restrict__ count, int E, const double *__restrict__ h_dev, const MtPeers peers) {
    extern __shared__ double Hs[];  // [E*E]
    const int n_items = *count;
    if (blockIdx.x * kListWarps >= n_items || n_items > kShortListMax) return;
    for (int i = threadIdx.x; i < E * E; i += blockDim.x) Hs[i] = h_dev[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kListWarps + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * kListWarps;
    const bool valid = lane < E;
    double hmin = valid ? Hs[lane * E + lane] : 1e300;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) hmin = fmin(hmin, __shfl_xor_sync(0xffffffffu, hmin, s));

    for (int item = warp_global; item < n_items; item += n_warps) {
        const long long px = list[item];
        const double xu = valid ? (double)x[(long long)lane * ld_x + px] : 0.0;
        double xmax = fabs(xu);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, s));
        const double tol_x = 1e-12 * xmax, tol_y = 1e-12 * xmax / hmin;
        unsigned A = __ballot_sync(0xffffffffu, valid && xu < 0.0);
        double xs = xu;
        int ninf = E + 1, backup = 3;
        for (int iter = 0; iter < 4 * E + 8 && A != 0u; ++iter) {
            const int n = __popc(A);
            const int a = (lane < n) ? (int)__fns(A, 0, lane + 1) : 0;  // component of my system row
            double v = __shfl_sync(0xffffffffu, xu, a);
            // the elimination is unrolled to the smallest of four sizes that holds the zero set: with 32-wide loops a
            // three-component system (the common case here) issued the predicated-off instructions of all 496 pairs
            double z;
            if (n <= 4)
                z = gauss_jordan_row<4>(Hs, E, a, n, lane, v);
            else if (n <= 8)
                z = gauss_jordan_row<8>(Hs, E, a, n, lane, v);
            else if (n <= 16)
                z = gauss_jordan_row<16>(Hs, E, a, n, lane, v);
            else
                z = gauss_jordan_row<32>(Hs, E, a, n, lane, v);
            const bool active = (A >> lane) & 1u;
            xs = xu;
            for (int r = 0; r < n; ++r) {
                const int ar = __shfl_sync(0xffffffffu, a, r);
                const double zr = __shfl_sync(0xffffffffu, z, r);
                if (valid && !active) xs -= Hs[lane * E + ar] * zr;
            }
            const double zmine = __shfl_sync(0xffffffffu, z, __popc(A & ((1u << lane) - 1u)));
            if (active) xs = -zmine;
            const bool bad = valid && (active ? (xs < -tol_y) : (xs < -tol_x));
            const unsigned V = __ballot_sync(0xffffffffu, bad);
            const int nb = __popc(V);
            if (nb == 0) break;
            unsigned flip = V;
            if (nb < ninf) {
                ninf = nb;
                backup = 3;
            } else if (backup > 0) {
                --backup;
            } else {
                flip = 1u << (31 - __clz(V));
            }
            A ^= flip;
        }
        if (valid) {
            const bool active = (A >> lane) & 1u;
            const float out = active ? 0.f : (float)fmax(xs, 0.0);
            const long long at = (long long)lane * ld_x + px;
            x[at] = out;
            if (peers.n_owned > 0 && peers.owner_row[lane]) peers.owner_row[lane][px] = out;
            for (int k = 0; k < peers.n_peers; ++k) peers.peer[k][at] = out;
            if (peers.multicast)
                asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(peers.multicast + at), "f"(out) : "memory");
        }
    }
}

int dispatch_nnls(float *x, int64_t ld_x, int x_layout, int64_t n_px, int E, const double *g, const double *h, int *n_fixed,
                  const int *list, const int *list_count, int list_min, const MtPeers *peers_host, cudaStream_t stream) {
    MtPeers peers{};
    if (peers_host) {
        MT_REQUIRE(peers_host->n_peers >= 0 && peers_host->n_peers <= MT_MAX_PEERS, "non-negative solve: n_peers=%d",
                   peers_host->n_peers);
        peers = *peers_host;
    }
    MT_REQUIRE(x_layout == MT_X_ELEMENT_MAJOR ||
                   (x_layout == MT_X_PIXEL_MAJOR && ld_x >= ((E + 3) & ~3) && ld_x % 4 == 0 && ((uintptr_t)x & 15) == 0),
               "non-negative solve: x_layout=%d ld_x=%lld", x_layout, (long long)ld_x);
    const int pixel_major = x_layout == MT_X_PIXEL_MAJOR;
    if (E <= 16) return launch_nnls<16>(x, ld_x, pixel_major, n_px, E, g, h, n_fixed, list, list_count, list_min, peers, stream);
    if (E <= 32) return launch_nnls<32>(x, ld_x, pixel_major, n_px, E, g, h, n_fixed, list, list_count, list_min, peers, stream);
    if (E <= 64) return launch_nnls<64>(x, ld_x, pixel_major, n_px, E, g, h, n_fixed, list, list_count, list_min, peers, stream);
    if (E <= 104) return launch_nnls<104>(x, ld_x, pixel_major, n_px, E, g, h, n_fixed, list, list_count, list_min, peers, stream);
    return mt::fail(MT_ERR_UNSUPPORTED, "non-negative solve: n_comp=%d > 104", E);
}

}  // namespace

extern "C" int mt_nnls_singular_count(uint64_t *count_host, int32_t reset) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(count_host, "mt_nnls_singular_count: null pointer");
    unsigned long long v = 0ull;
    MT_CUDA_TRY(cudaMemcpyFromSymbol(&v, g_singular_pixels, sizeof(v)));
    if (reset) {
        const unsigned long long zero = 0ull;
        MT_CUDA_TRY(cudaMemcpyToSymbol(g_singular_pixels, &zero, sizeof(zero)));
    }
    *count_host = v;
    return MT_OK;
}

extern "C" int mt_nnls_fixup(float *x_dev, int64_t ld_x, int32_t x_layout, int64_t n_px, int32_t n_comp,
                             const double *g_dev,
                             const double *hinv_dev, int32_t *n_fixed_dev, const MtPeers *peers_host, void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(x_dev && g_dev && hinv_dev, "mt_nnls_fixup: null pointer");
    MT_REQUIRE(n_px >= 0 && n_comp >= 1 && (x_layout == MT_X_PIXEL_MAJOR || ld_x >= n_px), "mt_nnls_fixup: bad sizes");
    if (n_px == 0) return MT_OK;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    return dispatch_nnls(x_dev, ld_x, x_layout, n_px, n_comp, g_dev, hinv_dev, n_fixed_dev, nullptr, nullptr, -1, peers_host,
                         stream);
}

extern "C" int mt_nnls_list(float *x_dev, int64_t ld_x, int32_t x_layout, const int32_t *list_dev,
                            const int32_t *count_dev, int32_t short_list,
                            int32_t n_comp, const double *g_dev, const double *hinv_dev, const MtPeers *peers_host,
                            void *stream_) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(x_dev && list_dev && count_dev && g_dev && hinv_dev, "mt_nnls_list: null pointer");
    MT_REQUIRE(n_comp >= 1, "mt_nnls_list: n_comp=%d", n_comp);
    if (short_list && n_comp <= 32 && x_layout == MT_X_ELEMENT_MAJOR) {
        MtPeers peers{};
        if (peers_host) peers = *peers_host;
        cudaStream_t stream = static_cast<cudaStream_t>(stream_);
        const size_t smem = sizeof(double) * (size_t)n_comp * n_comp;
        nnls_warp_kernel<<<mt::num_sms() * 8, kListWarps * 32, smem, stream>>>(x_dev, ld_x, list_dev, count_dev, n_comp,
                                                                             hinv_dev, peers);
        MT_CUDA_TRY(cudaGetLastError());
        // long lists: the thread-per-pixel kernel takes over (returns immediately for short ones)
        return dispatch_nnls(x_dev, ld_x, x_layout, 0, n_comp, g_dev, hinv_dev, nullptr, list_dev, count_dev, kShortListMax,
                             peers_host, stream);
    }
    return dispatch_nnls(x_dev, ld_x, x_layout, 0, n_comp, g_dev, hinv_dev, nullptr, list_dev, count_dev, -1, peers_host,
                         static_cast<cudaStream_t>(stream_));
}

