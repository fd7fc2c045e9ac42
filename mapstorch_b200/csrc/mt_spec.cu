// One-spectrum amplitude solve of fit_spec -- the evaluation inside the Levenberg-Marquardt loop of
// fit_spec(tune_params=True) (the reference's objective, opt.py:329-398, with the calibration held at a trial point).
//
// For given calibration parameters the model is linear in the amplitudes A = 10**a >= 0 (opt.py:154, map.py:163), so the
// inner problem is
//     x = argmin_{x >= 0}  sum_k w_k (x . B[:, c_k] - t_k)^2 + 2 * shift * sum(x),      t = y - bkg,
// over the fitted channels c_k (`indices` of fit_spec), with optional weights (the re-weighted rounds of loss="l1") and
// the l1_lambda penalty (opt.py:372-396) as a shift of the right-hand side.  Until round 2 this was ~15 torch fp64 ops,
// torch.linalg.inv, mt_nnls_fixup and three host synchronisations per trial point; here it is ONE launch, one CTA per
// problem (a batch = the trial points of one forward-difference Jacobian):
//   A  Gram matrix and right-hand side in fp64 (4 x 4 register blocks per warp over the channel loop, fixed reduction
//      order), the target t staged in shared memory;
//   B  block principal pivoting (the rule of mt_nnls.cu) on the normal equations, every pivot step a CTA-wide
//      right-looking Cholesky of G_FF in shared memory + two triangular solves + the duals of the zero set;
//   C  residual vector, its weighted form (what Levenberg-Marquardt differences), the objective and the penalty entries.
#include <algorithm>

#include "mt_common.cuh"

namespace {

constexpr int kSpecThreads = 512;
constexpr int kSpecMaxComp = MT_SPEC_MAX_COMP;
constexpr int kSpecMaxCh = MT_SPEC_MAX_CH;

struct SpecArgs {
    const float *basis;
    long long ld_b, prob_stride_b;
    int n_comp, n_k;
    const float *y, *bkg;
    long long ld_bkg;
    const int *chan_idx;
    const double *weights;
    long long ld_w;
    double l1_lambda;
    int robust, n_elements;
    double *x, *res, *vec, *stats;
    double *partial;  // [n_prob][slices][E * E + E] Gram / right-hand-side partial sums (slices > 1)
    int *arrived;     // [n_prob] CTAs of the problem that have delivered their partial sums
    unsigned char *free_set;  // optional [n_prob][E], in: starting partition (1 free, 0 zero, other: no hint), out: final one
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA in a fixed order; every thread receives the total.  scratch: kSpecThreads / 32 doubles.
__device__ double block_sum(double v, double *scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kSpecThreads / 32; ++w) s += scratch[w];
    return s;
}

__global__ void __launch_bounds__(kSpecThreads) spec_solve_kernel(const SpecArgs a) {
    extern __shared__ double smem[];
    const int E = a.n_comp, K = a.n_k, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int prob = blockIdx.y, n_slice = gridDim.x, slice = blockIdx.x;
    double *G = smem;                 // [E][E] Gram matrix (weighted), symmetric by construction
    double *W = G + E * E;            // [E][E] Cholesky workspace of the free block
    double *rhs = W + E * E;          // [E]
    double *xs = rhs + E;             // primal (free set) / dual (zero set)
    double *xu = xs + E;              // unconstrained solution of the live set (first pivot step)
    double *v = xu + E;               // right-hand side / forward-substituted vector of the free block
    double *z = v + E;                // solution of the free block
    double *red = z + E;              // [kSpecThreads / 32] reduction scratch
    float *tm = reinterpret_cast<float *>(red + kSpecThreads / 32);  // [K] target y - bkg (an fp32 difference, as the host code forms it)
    int *idx = reinterpret_cast<int *>(tm + ((K + 1) & ~1));         // [E] members of the free set
    unsigned char *free_ = reinterpret_cast<unsigned char *>(idx + E);  // [E] 1 free, 0 zero set, 2 dead (empty column)
    __shared__ int s_n, s_last_bad, s_is_last;

    const float *B = a.basis + (long long)prob * a.prob_stride_b;
    const float *bkg = a.bkg ? a.bkg + (long long)prob * a.ld_bkg : nullptr;
    const double *wts = a.weights ? a.weights + (long long)prob * a.ld_w : nullptr;
    double *x_out = a.x + (long long)prob * E;
    double *res_out = a.res + (long long)prob * K;
    double *vec_out = a.vec + (long long)prob * (K + E);
    double *stats = a.stats + (long long)prob * 8;

    // ---- target and the penalty constant c = l1_lambda * loss(bkg, y) / n_elements (opt.py:372-381)
    double t2 = 0.0, t1 = 0.0;
    for (int k = tid; k < K; k += kSpecThreads) {
        const int c = a.chan_idx ? a.chan_idx[k] : k;
        const float t = a.y[c] - (bkg ? bkg[c] : 0.f);
        tm[k] = t;
        t2 += (double)t * (double)t;
        t1 += fabs((double)t);
    }
    t2 = block_sum(t2, red);
    t1 = block_sum(t1, red);
    const double pen_c = a.l1_lambda > 0.0 ? a.l1_lambda * (a.robust ? t1 : t2) / (double)max(a.n_elements, 1) : 0.0;
    const double shift = a.robust ? pen_c : 0.5 * pen_c;
    __syncthreads();

    // ---- phase A: G = (B w) B^T and rhs = (B w) t; row index E stands for the target
    {
        const int nb = (E + 1 + 3) / 4;  // 4-row blocks over the E + 1 rows
        const int n_blocks = nb * (nb + 1) / 2;
        // wide problems are cut into channel slices, one CTA each (grid.x): the partial sums go to global memory and the
        // last CTA to arrive adds them up in slice order (deterministic) and carries on alone
        const int per = ((K + n_slice - 1) / n_slice + 31) & ~31;
        const int k_lo = min(K, slice * per), k_hi = min(K, k_lo + per);
        double *mine = n_slice > 1 ? a.partial + ((long long)prob * n_slice + slice) * (E * E + E) : nullptr;
        for (int blk = warp; blk < n_blocks; blk += kSpecThreads / 32) {
            int bi = 0, rem = blk;  // upper triangle, row-major: (bi, bj >= bi)
            while (rem >= nb - bi) {
                rem -= nb - bi;
                ++bi;
            }
            const int bj = bi + rem, r0 = bi * 4, s0 = bj * 4;
            double acc[4][4];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
            for (int k = k_lo + lane; k < k_hi; k += 32) {
                const int c = a.chan_idx ? a.chan_idx[k] : k;
                const double w = wts ? wts[k] : 1.0;
                const double t = (double)tm[k];
                double ra[4], rb[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int r = r0 + p, s = s0 + p;
                    ra[p] = (r < E ? (double)B[(long long)r * a.ld_b + c] : (r == E ? t : 0.0)) * w;
                    rb[p] = s < E ? (double)B[(long long)s * a.ld_b + c] : (s == E ? t : 0.0);
                }
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[p][q] = fma(ra[p], rb[q], acc[p][q]);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double sum = warp_sum(acc[p][q]);
                    const int r = r0 + p, s = s0 + q;
                    if (lane == 0 && r < E && s >= r) {  // upper triangle only: both halves get the same number
                        if (mine != nullptr) {
                            if (s <= E) mine[s < E ? r * E + s : E * E + r] = sum;
                        } else if (s < E) {
                            G[r * E + s] = sum;
                            G[s * E + r] = sum;
                        } else if (s == E) {
                            rhs[r] = sum - shift;
                        }
                    }
                }
        }
    }
    __syncthreads();
    if (n_slice > 1) {
        if (tid == 0) {
            __threadfence();
            s_is_last = atomicAdd(a.arrived + prob, 1) == n_slice - 1;
        }
        __syncthreads();
        if (!s_is_last) return;
        __threadfence();
        const double *part = a.partial + (long long)prob * n_slice * (E * E + E);
        for (int i = tid; i < E * E + E; i += kSpecThreads) {
            const int r = i < E * E ? i / E : i - E * E, c = i < E * E ? i - r * E : E;
            if (c < r) continue;  // the slices wrote the upper triangle only
            double sum = 0.0;
            for (int q = 0; q < n_slice; ++q) sum += __ldcg(part + (long long)q * (E * E + E) + i);
            if (c < E) {
                G[r * E + c] = sum;
                G[c * E + r] = sum;
            } else {
                rhs[r] = sum - shift;
            }
        }
        __syncthreads();
    }

    // ---- phase B: block principal pivoting.  Dead components (empty columns) never enter.
    double gmax = 0.0;
    for (int e = 0; e < E; ++e) gmax = fmax(gmax, G[e * E + e]);
    // Starting partition: everything free, or the partition the caller hands in -- the final one of a neighbouring trial
    // point of the Levenberg-Marquardt loop, which is nearly always already right (a 93-component problem took 33 pivot
    // steps from scratch, ~0.1 ms each).
    unsigned char *hint = a.free_set ? a.free_set + (long long)prob * E : nullptr;
    const bool my_zero = tid < E && hint != nullptr && hint[tid] == 0;
    bool warm = __syncthreads_or(my_zero) != 0;
    if (tid < E) {
        free_[tid] = G[tid * E + tid] > 0.0 ? (my_zero ? 0 : 1) : 2;
        xs[tid] = 0.0;
        xu[tid] = 0.0;
    }
    __syncthreads();
    int ninf = E + 1, backup = 3, iters = 0, flag = 0;
    double tol_x = 0.0, tol_y = 0.0;
    const int max_iter = 8 * E + 16;
    for (int iter = 0; iter < max_iter; ++iter) {
        ++iters;
        if (tid == 0) {
            int n = 0;
            for (int e = 0; e < E; ++e)
                if (free_[e] == 1) idx[n++] = e;
            s_n = n;
            s_last_bad = -1;
        }
        __syncthreads();
        const int n = s_n;
        for (int t = tid; t < n * n; t += kSpecThreads) {
            const int i = t / n, j = t - i * n;
            if (j <= i) W[i * n + j] = G[idx[i] * E + idx[j]];
        }
        if (tid < n) v[tid] = rhs[idx[tid]];
        __syncthreads();
        // right-looking Cholesky, lower triangle in place
        bool ok = true;
        for (int j = 0; j < n; ++j) {
            // a pivot below 1e-13 of its diagonal entry is rounding noise of collinear columns, not information
            const double p = W[j * n + j];
            if (!(p > 1e-13 * G[idx[j] * E + idx[j]])) {
                ok = false;
                break;  // uniform: every thread read the same value
            }
            const double d = sqrt(p);
            __syncthreads();
            for (int i = j + tid; i < n; i += kSpecThreads) W[i * n + j] = (i == j) ? d : W[i * n + j] / d;
            __syncthreads();
            // trailing update of the lower triangle: a warp per row, lanes along the row
            for (int i = j + 1 + warp; i < n; i += kSpecThreads / 32) {
                const double lij = W[i * n + j];
                for (int k = j + 1 + lane; k <= i; k += 32) W[i * n + k] = fma(-lij, W[k * n + j], W[i * n + k]);
            }
            __syncthreads();
        }
        if (!ok && warm) {
            // the handed-in partition does not fit this problem: start over from the full live set
            __syncthreads();
            if (tid < E && free_[tid] != 2) free_[tid] = 1;
            warm = false;
            ninf = E + 1;
            backup = 3;
            iter = -1;
            __syncthreads();
            continue;
        }
        if (!ok) {
            // numerically singular subset (strongly overlapping components).  On the first step nothing has been solved
            // yet: the problem is reported as failed; later: the clamped unconstrained solution, as mt_nnls.cu does
            flag = iter == 0 ? 2 : 1;
            __syncthreads();
            if (tid < E && free_[tid] != 2) {
                free_[tid] = xu[tid] > 0.0 ? 1 : 0;
                xs[tid] = xu[tid];
            }
            __syncthreads();
            break;
        }
        for (int j = 0; j < n; ++j) {  // L f = v (column oriented): f -> z
            const double fj = v[j] / W[j * n + j];
            if (tid == 0) z[j] = fj;
            for (int i = j + 1 + tid; i < n; i += kSpecThreads) v[i] = fma(-W[i * n + j], fj, v[i]);
            __syncthreads();
        }
        for (int j = n - 1; j >= 0; --j) {  // L^T s = f: s -> v
            const double sj = z[j] / W[j * n + j];
            if (tid == 0) v[j] = sj;
            for (int i = tid; i < j; i += kSpecThreads) z[i] = fma(-W[j * n + i], sj, z[i]);
            __syncthreads();
        }
        if (tid < n) xs[idx[tid]] = v[tid];
        __syncthreads();
        if (tid < E && free_[tid] == 0) {  // duals of the zero set: (G x - rhs)_e
            double s = -rhs[tid];
            const double *ge = G + tid * E;
            for (int k = 0; k < n; ++k) s = fma(ge[idx[k]], v[k], s);
            xs[tid] = s;
        }
        __syncthreads();
        if (iter == 0) {
            double xmax = 0.0;  // scale of the amplitudes (free entries: with a handed-in partition the others hold duals)
            for (int e = 0; e < E; ++e)
                if (free_[e] == 1) xmax = fmax(xmax, fabs(xs[e]));
            tol_x = 1e-12 * xmax;
            tol_y = 1e-12 * xmax * gmax;
            if (tid < E) xu[tid] = xs[tid];
        }
        const bool bad = tid < E && free_[tid] != 2 && (free_[tid] == 1 ? (xs[tid] < -tol_x) : (xs[tid] < -tol_y));
        if (bad) atomicMax(&s_last_bad, tid);
        const int n_bad = __syncthreads_count(bad);
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
        if (bad && (flip_all || tid == s_last_bad)) free_[tid] = free_[tid] == 1 ? 0 : 1;
        __syncthreads();
    }
    if (tid < E && hint != nullptr) hint[tid] = free_[tid];
    if (tid < E) {
        const double out = (free_[tid] == 1 && flag != 2) ? fmax(xs[tid], 0.0) : 0.0;
        xs[tid] = out;
        x_out[tid] = out;
    }
    __syncthreads();

    // ---- phase C: residual, weighted residual, objective, penalty entries
    double sv = 0.0, s2 = 0.0, s1 = 0.0;
    for (int k = tid; k < K; k += kSpecThreads) {
        const int c = a.chan_idx ? a.chan_idx[k] : k;
        double m = 0.0;
        for (int e = 0; e < E; ++e) m = fma(xs[e], (double)B[(long long)e * a.ld_b + c], m);
        const double r = m - (double)tm[k];
        const double wr = wts ? sqrt(wts[k]) * r : r;
        res_out[k] = r;
        vec_out[k] = wr;
        sv = fma(wr, wr, sv);
        s2 = fma(r, r, s2);
        s1 += fabs(r);
    }
    sv = block_sum(sv, red);
    s2 = block_sum(s2, red);
    s1 = block_sum(s1, red);
    double xsum = 0.0;
    for (int e = 0; e < E; ++e) xsum += xs[e];
    const double pk = a.robust ? 2.0 : 1.0;  // the penalty enters the residual vector as sqrt(pk * c * x) entries
    if (tid < E) vec_out[K + tid] = sqrt(pk * pen_c * xs[tid]);
    if (tid == 0) {
        stats[0] = sv + pk * pen_c * xsum;                     // |vec|^2, what Levenberg-Marquardt compares
        stats[1] = (a.robust ? s1 : s2) + pen_c * xsum;        // the reference's objective
        stats[2] = pen_c;
        stats[3] = (double)flag;
        stats[4] = (double)iters;
        stats[5] = s2;
        stats[6] = s1;
        stats[7] = xsum;
    }
}

}  // namespace

extern "C" int mt_spec_solve(const float *basis_dev, int64_t ld_b, int64_t prob_stride_b, int32_t n_comp, const float *y_dev,
                             const float *bkg_dev, int64_t ld_bkg, const int32_t *chan_idx_dev, int32_t n_k,
                             const double *weights_dev, int64_t ld_w, double l1_lambda, int32_t robust, int32_t n_elements,
                             double *x_dev, double *res_dev, double *vec_dev, double *stats_dev, uint8_t *free_set_dev,
                             int32_t n_prob, void *stream) {
    if (int rc = mt::require_sm100()) return rc;
    MT_REQUIRE(basis_dev && y_dev && x_dev && res_dev && vec_dev && stats_dev, "mt_spec_solve: null pointer");
    MT_REQUIRE(n_comp >= 1 && n_comp <= kSpecMaxComp, "mt_spec_solve: n_comp %d outside [1, %d]", n_comp, kSpecMaxComp);
    MT_REQUIRE(n_k >= 1 && n_k <= kSpecMaxCh, "mt_spec_solve: %d fitted channels outside [1, %d]", n_k, kSpecMaxCh);
    MT_REQUIRE(n_prob >= 1 && n_prob <= 65535, "mt_spec_solve: n_prob %d", n_prob);  // grid.y
    MT_REQUIRE(ld_b >= 1 && ld_bkg >= 0 && ld_w >= 0, "mt_spec_solve: negative pitch");
    SpecArgs a;
    a.basis = basis_dev;
    a.ld_b = ld_b;
    a.prob_stride_b = prob_stride_b;
    a.n_comp = n_comp;
    a.n_k = n_k;
    a.y = y_dev;
    a.bkg = bkg_dev;
    a.ld_bkg = ld_bkg;
    a.chan_idx = chan_idx_dev;
    a.weights = weights_dev;
    a.ld_w = ld_w;
    a.l1_lambda = l1_lambda;
    a.robust = robust;
    a.n_elements = n_elements;
    a.x = x_dev;
    a.res = res_dev;
    a.vec = vec_dev;
    a.stats = stats_dev;
    a.free_set = free_set_dev;
    const size_t E = (size_t)n_comp;
    const size_t smem = sizeof(double) * (2 * E * E + 5 * E + kSpecThreads / 32) + sizeof(float) * (((size_t)n_k + 1) & ~(size_t)1) +
                        sizeof(int) * E + ((E + 7) & ~(size_t)7);
    static int smem_set[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem_set[dev] < (int)smem) {
        MT_CUDA_TRY(cudaFuncSetAttribute(spec_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) smem_set[dev] = (int)smem;
    }
    // channel slices per problem: one per ~20 k (4 x 4 block, channel) products of the Gram phase, at most 32
    const long long nb = (n_comp + 1 + 3) / 4, work = nb * (nb + 1) / 2 * (long long)n_k;
    int n_slice = (int)std::min<long long>(32, std::max<long long>(1, work / 20000));
    n_slice = std::min(n_slice, std::max(1, n_k / 64));
    mt::StreamScratch scratch;
    a.partial = nullptr;
    a.arrived = nullptr;
    if (n_slice > 1) {
        const size_t part_bytes = sizeof(double) * (size_t)n_prob * n_slice * (E * E + E);
        const size_t count_bytes = (sizeof(int) * (size_t)n_prob + 15) & ~(size_t)15;
        MT_CUDA_TRY(scratch.alloc(part_bytes + count_bytes, static_cast<cudaStream_t>(stream)));
        a.partial = reinterpret_cast<double *>(scratch.ptr);
        a.arrived = reinterpret_cast<int *>(scratch.ptr + part_bytes);
        MT_CUDA_TRY(cudaMemsetAsync(a.arrived, 0, count_bytes, static_cast<cudaStream_t>(stream)));
    }
    spec_solve_kernel<<<dim3((unsigned)n_slice, (unsigned)n_prob), kSpecThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    MT_CUDA_TRY(cudaGetLastError());
    return MT_OK;
}
