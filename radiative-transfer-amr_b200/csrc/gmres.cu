// Restarted GMRES on the device for A_ii x = b (replaces KSPSolve with
// KSPGMRES / KSPFGMRES, restart 50: src/rt/solve.py:80-100; the arithmetic of
// the reference lives in PETSc 3.22, which is not part of the reference tree).
//
// Arnoldi with classical Gram-Schmidt in two fused, bandwidth-bound kernels:
//   k_mdot : h[i] = <V_i, w>, i < k, and <w, w>; w is read ONCE, every basis
//            vector once (accumulators for all k dot products live in registers)
//   k_maxpy: v = w - sum_i h[i] V_i written straight into the next basis slot,
//            with ||v||^2 accumulated on the fly
// followed by a scaling pass.  Reductions are two-stage (per-block partials,
// then one block per output) so results are deterministic.  The small
// Hessenberg least-squares problem (Givens rotations) runs on the host, ONE ITERATION BEHIND
// the device: the k+2 scalars of iteration k travel to pinned host memory with an
// asynchronous copy + event, and the host reads them after it has enqueued iteration k+1,
// so the stream never drains between iterations (a converged solve has enqueued one
// Arnoldi step more than it uses).
//
// Preconditioning (rtb200_gmres_params.pc / pc_side): none; left (PETSc's KSPGMRES default,
// residual norms of the preconditioned system); right and flexible (KSPFGMRES, which both
// scatterer experiments of the reference select, solve.py:83 / input.json:4): the
// z_k = M^-1 v_k are stored, x is updated with them and the residual norms are those of the
// original system.
#include "rtb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

constexpr int RED_BLOCKS_PER_SM = 4;
constexpr int RED_THREADS = 256;
constexpr int MAX_BASIS = 1024;       // restart + 1 must not exceed this
constexpr int RED_STRIDE = 68;        // doubles per block in the partial-sum buffer (<= 64 outputs per pass)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// partial[blockIdx.x * stride + i] = sum over this block's elements of V_i * w,
// i < k; partial[... + k] = sum w * w.   KB >= k + 1 is the unrolled bound.
template <int KB>
__global__ void __launch_bounds__(RED_THREADS)
k_mdot(const double* __restrict__ V, int64_t ld, int k, const double* __restrict__ w,
       int64_t n2 /* n / 2 */, double* __restrict__ partial, int stride) {
    double acc[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) acc[i] = 0.0;
    const double2* __restrict__ w2 = reinterpret_cast<const double2*>(w);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += step) {
        const double2 wv = w2[e];
#pragma unroll
        for (int i = 0; i < KB - 1; ++i) {
            if (i < k) {
                const double2 vv = reinterpret_cast<const double2*>(V + (int64_t)i * ld)[e];
                acc[i] = fma(vv.x, wv.x, fma(vv.y, wv.y, acc[i]));
            }
        }
        acc[KB - 1] = fma(wv.x, wv.x, fma(wv.y, wv.y, acc[KB - 1]));
    }
    __shared__ double red[RED_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        if (i < k || i == KB - 1) {
            double s = warp_sum(acc[i]);
            if (lane == 0) red[wid] = s;
            __syncthreads();
            if (wid == 0) {
                s = lane < RED_THREADS / 32 ? red[lane] : 0.0;
                s = warp_sum(s);
                if (lane == 0) partial[(int64_t)blockIdx.x * stride + (i == KB - 1 ? k : i)] = s;
            }
            __syncthreads();
        }
    }
}

// out[i] (+)= sum_b partial[b * stride + i]   (one block per output)
__global__ void __launch_bounds__(RED_THREADS)
k_reduce(const double* __restrict__ partial, int nblocks, int stride, double* __restrict__ out,
         int accumulate) {
    const int i = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[(int64_t)b * stride + i];
    __shared__ double red[RED_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (wid == 0) {
        s = lane < RED_THREADS / 32 ? red[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) out[i] = accumulate ? out[i] + s : s;
    }
}

// out = w + sign * sum_i h[i] V_i ; partial[blockIdx.x] = sum out^2
// skip_flag != nullptr and *skip_flag == 0: do nothing (conditional refinement).
__global__ void __launch_bounds__(RED_THREADS)
k_maxpy(const double* __restrict__ V, int64_t ld, int k, const double* __restrict__ h,
        double sign, const double* w, double* out, int64_t n2, double* __restrict__ partial,
        const int* __restrict__ skip_flag) {
    if (skip_flag && *skip_flag == 0) return;
    __shared__ double hs[MAX_BASIS];
    for (int i = threadIdx.x; i < k; i += blockDim.x) hs[i] = sign * h[i];
    __syncthreads();
    const double2* w2 = reinterpret_cast<const double2*>(w);
    double2* o2 = reinterpret_cast<double2*>(out);
    double nrm = 0.0;
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += step) {
        double2 a = w2[e];
#pragma unroll 4
        for (int i = 0; i < k; ++i) {
            const double2 vv = reinterpret_cast<const double2*>(V + (int64_t)i * ld)[e];
            a.x = fma(hs[i], vv.x, a.x);
            a.y = fma(hs[i], vv.y, a.y);
        }
        o2[e] = a;
        nrm = fma(a.x, a.x, fma(a.y, a.y, nrm));
    }
    if (partial) {
        __shared__ double red[RED_THREADS / 32];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        nrm = warp_sum(nrm);
        if (lane == 0) red[wid] = nrm;
        __syncthreads();
        if (wid == 0) {
            nrm = lane < RED_THREADS / 32 ? red[lane] : 0.0;
            nrm = warp_sum(nrm);
            if (lane == 0) partial[blockIdx.x] = nrm;
        }
    }
}

// v *= 1 / sqrt(*nrm2)   (no-op when the norm is zero: breakdown is handled on the host)
__global__ void k_scale_inv_norm(double* v, int64_t n2, const double* __restrict__ nrm2) {
    const double s = *nrm2;
    if (!(s > 0.0)) return;
    const double f = 1.0 / sqrt(s);
    double2* v2 = reinterpret_cast<double2*>(v);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += step) {
        double2 a = v2[e];
        a.x *= f;
        a.y *= f;
        v2[e] = a;
    }
}

// out = a - b
__global__ void k_sub(const double* __restrict__ a, const double* __restrict__ b, double* out,
                      int64_t n2) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
    double2* o2 = reinterpret_cast<double2*>(out);
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n2; e += step) {
        double2 x = a2[e], y = b2[e];
        x.x -= y.x;
        x.y -= y.y;
        o2[e] = x;
    }
}

// DGKS test: flag = (new_norm2 < 0.5 * old_norm2), i.e. more than 1/sqrt(2) cancelled
__global__ void k_dgks_flag(const double* nrm2_new, const double* nrm2_old, int* flag) {
    *flag = (*nrm2_new < 0.5 * *nrm2_old) ? 1 : 0;
}

// h[i] += h2[i] for i < k when flag set
__global__ void k_add_h(double* h, const double* h2, int k, const int* flag) {
    if (flag && *flag == 0) return;
    for (int i = threadIdx.x; i < k; i += blockDim.x) h[i] += h2[i];
}

struct Work {
    double* V;          // (m + 1) basis vectors
    double* w;          // operator output
    double* t;          // scratch vector (preconditioner / residual)
    double* partial;    // reduction partials
    double* h;          // [MAX_BASIS + 4] dots; h[k] = <w,w>
    double* h2;         // refinement dots
    double* nrm;        // [4] squared norms
    int* flag;
    double* Z;          // m preconditioned basis vectors (flexible right preconditioning) or null
    int nblocks;
    int stride;
};

struct Solver {
    rtb200_ctx* ctx;
    cudaStream_t st;
    int64_t n, n2;
    Work wk;

    int mdot(const double* V, int k, const double* w, double* out) {
        // out[0..k-1] = V^T w, out[k] = w^T w
        const int kb = k + 1;
        double* p = wk.partial;
        const int nb = wk.nblocks, sd = wk.stride;
#define MD(KB) k_mdot<KB><<<nb, RED_THREADS, 0, st>>>(V, n, k, w, n2, p, sd)
        if (kb <= 2) MD(2);
        else if (kb <= 4) MD(4);
        else if (kb <= 8) MD(8);
        else if (kb <= 16) MD(16);
        else if (kb <= 24) MD(24);
        else if (kb <= 32) MD(32);
        else if (kb <= 40) MD(40);
        else if (kb <= 52) MD(52);
        else if (kb <= 64) MD(64);
        else {
            // very long restarts: chunks of 63 basis vectors (w re-read per chunk)
            for (int k0 = 0; k0 < k; k0 += 63) {
                const int kc = std::min(63, k - k0);
                k_mdot<64><<<nb, RED_THREADS, 0, st>>>(V + (int64_t)k0 * n, n, kc, w, n2, p, sd);
                k_reduce<<<kc, RED_THREADS, 0, st>>>(p, nb, sd, out + k0, 0);
                if (k0 + kc >= k) k_reduce<<<1, RED_THREADS, 0, st>>>(p + kc, nb, sd, out + k, 0);
                ctx->launches += 2;
            }
            ctx->launches += 1;
            return finish_reduce(out, k + 1);
        }
#undef MD
        k_reduce<<<k + 1, RED_THREADS, 0, st>>>(p, nb, sd, out, 0);
        ctx->launches += 2;
        return finish_reduce(out, k + 1);
    }

    int finish_reduce(double* out, int count) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rt_cuda(ctx, e, "gmres reduction launch");
        if (ctx->nranks > 1) return rt_allreduce_sum(ctx, out, count, st);
        return RTB200_OK;
    }

    int maxpy(const double* V, int k, const double* h, double sign, const double* w, double* out,
              double* nrm2_out, const int* flag) {
        k_maxpy<<<wk.nblocks, RED_THREADS, 0, st>>>(V, n, k, h, sign, w, out, n2,
                                                     nrm2_out ? wk.partial : nullptr, flag);
        ctx->launches++;
        if (nrm2_out) {
            // when the refinement is skipped the partials are stale: only reduce if flag set
            if (flag) return RTB200_OK;   // caller handles conditional norm
            k_reduce<<<1, RED_THREADS, 0, st>>>(wk.partial, wk.nblocks, 1, nrm2_out, 0);
            ctx->launches++;
            return finish_reduce(nrm2_out, 1);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return rt_cuda(ctx, e, "k_maxpy launch");
        return RTB200_OK;
    }

    int norm2(const double* v, double* out_dev) {
        // squared norm through the mdot kernel with k = 0
        k_mdot<2><<<wk.nblocks, RED_THREADS, 0, st>>>(v, n, 0, v, n2, wk.partial, wk.stride);
        k_reduce<<<1, RED_THREADS, 0, st>>>(wk.partial, wk.nblocks, wk.stride, out_dev, 0);
        ctx->launches += 2;
        return finish_reduce(out_dev, 1);
    }

    int fetch(const double* dev, double* host, int count) {
        cudaError_t e = cudaMemcpyAsync(host, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaMemcpyAsync D2H");
        e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaStreamSynchronize");
        return RTB200_OK;
    }

    // w = M^-1 A v (left), w = A M^-1 v with z = M^-1 v kept in zk (right, flexible), or w = A v
    int op(const double* v, double* w, int pc, int side, double* zk) {
        int rc;
        if (pc && side == 0) {
            rc = rt_apply_with_halo(ctx, RTB200_TERM_ALL, 1, v, wk.t, st);
            if (rc) return rc;
            return rt_pc_apply_type(ctx, pc, wk.t, w, st);
        }
        if (pc) {
            rc = rt_pc_apply_type(ctx, pc, v, zk, st);
            if (rc) return rc;
            return rt_apply_with_halo(ctx, RTB200_TERM_ALL, 1, zk, w, st);
        }
        return rt_apply_with_halo(ctx, RTB200_TERM_ALL, 1, v, w, st);
    }
};

// singular values of the (rows x cols) column-major matrix H (one-sided Jacobi)
void extreme_singular_values(std::vector<double> H, int rows, int cols, double* smax, double* smin) {
    *smax = 0.0;
    *smin = 0.0;
    if (cols <= 0) return;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < cols - 1; ++p)
            for (int q = p + 1; q < cols; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < rows; ++i) {
                    const double x = H[(size_t)p * rows + i], y = H[(size_t)q * rows + i];
                    a += x * x; b += y * y; g += x * y;
                }
                if (std::fabs(g) <= 1e-300 || std::fabs(g) <= 1e-15 * std::sqrt(a * b)) continue;
                off = std::max(off, std::fabs(g) / std::sqrt(a * b));
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < rows; ++i) {
                    const double x = H[(size_t)p * rows + i], y = H[(size_t)q * rows + i];
                    H[(size_t)p * rows + i] = cs * x - sn * y;
                    H[(size_t)q * rows + i] = sn * x + cs * y;
                }
            }
        if (off < 1e-14) break;
    }
    double mx = 0.0, mn = 1e300;
    for (int p = 0; p < cols; ++p) {
        double a = 0;
        for (int i = 0; i < rows; ++i) a += H[(size_t)p * rows + i] * H[(size_t)p * rows + i];
        a = std::sqrt(a);
        mx = std::max(mx, a);
        mn = std::min(mn, a);
    }
    *smax = mx;
    *smin = mn;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" int64_t rtb200_gmres_work_bytes(const rtb200_ctx* ctx, int restart, int pc_side) {
    if (!ctx || !ctx->have_mesh || restart < 1 || restart + 1 > MAX_BASIS) return -1;
    const size_t n = (size_t)ctx->dm.n_slab;
    const int nblocks = ctx->sm_count * RED_BLOCKS_PER_SM;
    size_t bytes = 0;
    bytes += align_up(sizeof(double) * n * (size_t)(restart + 1), 256);   // V
    if (pc_side == 1) bytes += align_up(sizeof(double) * n * (size_t)restart, 256);   // Z
    bytes += align_up(sizeof(double) * n, 256) * 2;                        // w, t
    bytes += align_up(sizeof(double) * (size_t)nblocks * RED_STRIDE, 256);
    bytes += align_up(sizeof(double) * (MAX_BASIS + 4), 256) * 2;          // h, h2
    bytes += 256 * 2;                                                      // nrm, flag
    return (int64_t)bytes;
}

extern "C" int rtb200_gmres(rtb200_ctx* ctx, const double* b_dev, double* x_dev,
                            const rtb200_gmres_params* prm, rtb200_gmres_result* res,
                            double* hist, void* work_dev, int64_t work_bytes, void* stream) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set");
    if (!b_dev || !x_dev || !prm || !res || !work_dev)
        return rt_fail(ctx, RTB200_EINVAL, "gmres: null argument");
    const int m = prm->restart;
    if (m < 1 || m + 1 > MAX_BASIS) return rt_fail(ctx, RTB200_EUNSUPPORTED, "gmres: restart out of range");
    const int pc = prm->pc;
    if (pc < RTB200_PC_NONE || pc > RTB200_PC_ELEMENT) return rt_fail(ctx, RTB200_EINVAL, "gmres: unknown preconditioner");
    if (prm->pc_side != 0 && prm->pc_side != 1) return rt_fail(ctx, RTB200_EINVAL, "gmres: pc_side must be 0 or 1");
    const int side = pc ? prm->pc_side : 0;
    if (work_bytes < rtb200_gmres_work_bytes(ctx, m, side))
        return rt_fail(ctx, RTB200_EINVAL, "gmres: workspace too small");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));

    Solver S;
    S.ctx = ctx;
    S.st = (cudaStream_t)stream;
    S.n = ctx->dm.n_slab;
    S.n2 = S.n / 2;
    const size_t n = (size_t)S.n;
    char* p = (char*)work_dev;
    Work& W = S.wk;
    W.nblocks = ctx->sm_count * RED_BLOCKS_PER_SM;
    W.stride = RED_STRIDE;
    W.V = (double*)p; p += align_up(sizeof(double) * n * (size_t)(m + 1), 256);
    W.Z = nullptr;
    if (side == 1) { W.Z = (double*)p; p += align_up(sizeof(double) * n * (size_t)m, 256); }
    W.w = (double*)p; p += align_up(sizeof(double) * n, 256);
    W.t = (double*)p; p += align_up(sizeof(double) * n, 256);
    W.partial = (double*)p; p += align_up(sizeof(double) * (size_t)W.nblocks * W.stride, 256);
    W.h = (double*)p; p += align_up(sizeof(double) * (MAX_BASIS + 4), 256);
    W.h2 = (double*)p; p += align_up(sizeof(double) * (MAX_BASIS + 4), 256);
    W.nrm = (double*)p; p += 256;
    W.flag = (int*)p; p += 256;

    // pinned staging for the per-iteration scalars: two slots, the host runs one iteration behind
    struct Pinned {
        double* buf = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        ~Pinned() {
            if (buf) cudaFreeHost(buf);
            for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
        }
    } pin;
    RT_CUDA(ctx, cudaMallocHost((void**)&pin.buf, sizeof(double) * 2 * (MAX_BASIS + 8)));
    for (int i = 0; i < 2; ++i) RT_CUDA(ctx, cudaEventCreateWithFlags(&pin.ev[i], cudaEventDisableTiming));

    if (pc == RTB200_PC_NODE) {
        int rc = rt_pc_setup(ctx, S.st);
        if (rc) return rc;
    }
    int64_t n_apply = 0;
    int rc;
    double host[8];

    // ||b|| (preconditioned when a left preconditioner is used, like PETSc)
    const double* bb = b_dev;
    if (pc && side == 0) {
        rc = rt_pc_apply_type(ctx, pc, b_dev, W.t, S.st);
        if (rc) return rc;
        bb = W.t;
    }
    rc = S.norm2(bb, W.nrm);
    if (rc) return rc;
    rc = S.fetch(W.nrm, host, 1);
    if (rc) return rc;
    const double bnorm = std::sqrt(host[0]);
    const double ttol = std::max(prm->rtol * bnorm, prm->atol);

    // Hessenberg data on the host (column-major, (m+1) x m), rotations, rhs g
    std::vector<double> H((size_t)(m + 1) * m, 0.0), Hraw((size_t)(m + 1) * m, 0.0);
    std::vector<double> cs(m + 1, 0.0), sn(m + 1, 0.0), g(m + 2, 0.0), yv(m + 1, 0.0);

    int it = 0, reason = 0;
    double resn = 0.0, res0 = 0.0;
    int last_cols = 0;
    bool first = true;

    while (reason == 0) {
        // r = b - A x   (x = 0: r = b)
        if (first && !prm->x0_nonzero) {
            RT_CUDA(ctx, cudaMemsetAsync(x_dev, 0, sizeof(double) * n, S.st));
            RT_CUDA(ctx, cudaMemcpyAsync(W.V, b_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, S.st));
        } else {
            rc = rt_apply_with_halo(ctx, RTB200_TERM_ALL, 1, x_dev, W.w, S.st);
            if (rc) return rc;
            n_apply++;
            k_sub<<<W.nblocks, RED_THREADS, 0, S.st>>>(b_dev, W.w, W.V, S.n2);
            ctx->launches++;
        }
        if (pc && side == 0) {
            rc = rt_pc_apply_type(ctx, pc, W.V, W.t, S.st);
            if (rc) return rc;
            RT_CUDA(ctx, cudaMemcpyAsync(W.V, W.t, sizeof(double) * n, cudaMemcpyDeviceToDevice, S.st));
        }
        rc = S.norm2(W.V, W.nrm);
        if (rc) return rc;
        rc = S.fetch(W.nrm, host, 1);
        if (rc) return rc;
        const double beta = std::sqrt(host[0]);
        if (first) {
            res0 = beta;
            resn = beta;
            if (hist && prm->record_history) hist[0] = beta;
            first = false;
            if (beta <= ttol) { reason = beta < prm->atol ? 3 : 2; break; }
        }
        if (!(beta > 0.0)) { reason = 3; resn = 0.0; break; }
        k_scale_inv_norm<<<W.nblocks, RED_THREADS, 0, S.st>>>(W.V, S.n2, W.nrm);
        ctx->launches++;
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        std::fill(H.begin(), H.end(), 0.0);
        std::fill(Hraw.begin(), Hraw.end(), 0.0);

        // device side of Arnoldi step k; its k + 2 scalars land in pinned slot k & 1
        auto enqueue = [&](int k) -> int {
            double* vk1 = W.V + (size_t)(k + 1) * n;
            int r2 = S.op(W.V + (size_t)k * n, W.w, pc, side, W.Z ? W.Z + (size_t)k * n : nullptr);
            if (r2) return r2;
            n_apply++;
            r2 = S.mdot(W.V, k + 1, W.w, W.h);                        // h = V^T w, h[k+1] = ||w||^2
            if (r2) return r2;
            r2 = S.maxpy(W.V, k + 1, W.h, -1.0, W.w, vk1, W.nrm, nullptr);   // v_{k+1} = w - V h
            if (r2) return r2;
            if (prm->orth != 2) {
                // refinement pass (always, or when more than 1/sqrt(2) of w cancelled)
                const int* flag = nullptr;
                if (prm->orth == 0) {
                    k_dgks_flag<<<1, 1, 0, S.st>>>(W.nrm, W.h + (k + 1), W.flag);
                    ctx->launches++;
                    flag = W.flag;
                }
                r2 = S.mdot(W.V, k + 1, vk1, W.h2);
                if (r2) return r2;
                k_maxpy<<<W.nblocks, RED_THREADS, 0, S.st>>>(W.V, S.n, k + 1, W.h2, -1.0, vk1, vk1,
                                                             S.n2, nullptr, flag);
                k_add_h<<<1, 64, 0, S.st>>>(W.h, W.h2, k + 1, flag);
                ctx->launches += 2;
                r2 = S.norm2(vk1, W.nrm);
                if (r2) return r2;
            }
            k_scale_inv_norm<<<W.nblocks, RED_THREADS, 0, S.st>>>(vk1, S.n2, W.nrm);
            ctx->launches++;
            double* slot = pin.buf + (size_t)(k & 1) * (MAX_BASIS + 8);
            RT_CUDA(ctx, cudaMemcpyAsync(slot, W.h, sizeof(double) * (k + 1), cudaMemcpyDeviceToHost, S.st));
            RT_CUDA(ctx, cudaMemcpyAsync(slot + (k + 1), W.nrm, sizeof(double), cudaMemcpyDeviceToHost, S.st));
            RT_CUDA(ctx, cudaEventRecord(pin.ev[k & 1], S.st));
            return RTB200_OK;
        };
        // host side of step k: Hessenberg column, rotations, convergence test
        auto process = [&](int k) -> int {
            RT_CUDA(ctx, cudaEventSynchronize(pin.ev[k & 1]));
            const double* hk = pin.buf + (size_t)(k & 1) * (MAX_BASIS + 8);
            const double hk1 = std::sqrt(std::max(hk[k + 1], 0.0));
            double* Hc = &H[(size_t)k * (m + 1)];
            for (int i = 0; i <= k; ++i) Hc[i] = hk[i];
            Hc[k + 1] = hk1;
            std::memcpy(&Hraw[(size_t)k * (m + 1)], Hc, sizeof(double) * (k + 2));
            for (int i = 0; i < k; ++i) {                       // previous rotations, then the new one
                const double a = cs[i] * Hc[i] + sn[i] * Hc[i + 1];
                Hc[i + 1] = -sn[i] * Hc[i] + cs[i] * Hc[i + 1];
                Hc[i] = a;
            }
            const double den = std::hypot(Hc[k], Hc[k + 1]);
            if (den == 0.0) { reason = -5; return RTB200_OK; }  // column k is not used
            cs[k] = Hc[k] / den;
            sn[k] = Hc[k + 1] / den;
            Hc[k] = den;
            Hc[k + 1] = 0.0;
            g[k + 1] = -sn[k] * g[k];
            g[k] = cs[k] * g[k];
            resn = std::fabs(g[k + 1]);
            ++it;
            last_cols = k + 1;
            if (hist && prm->record_history) hist[it] = resn;
            if (resn != resn) reason = -9;                       // NaN
            else if (resn <= ttol) reason = resn < prm->atol ? 3 : 2;
            else if (resn >= prm->divtol * res0) reason = -4;
            else if (it >= prm->max_it) reason = -3;
            else if (hk1 == 0.0) reason = -5;                    // happy breakdown without convergence
            return RTB200_OK;
        };

        last_cols = 0;
        int enq = 0;                                             // steps enqueued in this cycle
        for (int k = 0; k < m && reason == 0; ++k) {
            rc = enqueue(k);
            if (rc) return rc;
            enq = k + 1;
            if (k >= 1) {
                rc = process(k - 1);
                if (rc) return rc;
            }
        }
        if (reason == 0 && enq > 0) {
            rc = process(enq - 1);
            if (rc) return rc;
        }
        // last_cols columns are valid: solve the triangular system and update x
        const int kc = last_cols;
        for (int i = kc - 1; i >= 0; --i) {
            double s = g[i];
            for (int j = i + 1; j < kc; ++j) s -= H[(size_t)j * (m + 1) + i] * yv[j];
            yv[i] = s / H[(size_t)i * (m + 1) + i];
        }
        if (kc > 0) {
            RT_CUDA(ctx, cudaMemcpyAsync(W.h, yv.data(), sizeof(double) * kc, cudaMemcpyHostToDevice, S.st));
            rc = S.maxpy(W.Z ? W.Z : W.V, kc, W.h, +1.0, x_dev, x_dev, nullptr, nullptr);
            if (rc) return rc;
            RT_CUDA(ctx, cudaStreamSynchronize(S.st));    // yv is reused by the next cycle
        } else if (reason == 0) {
            reason = -5;                                  // no usable direction: do not spin
        }
    }

    // true residual at exit
    rc = rt_apply_with_halo(ctx, RTB200_TERM_ALL, 1, x_dev, W.w, S.st);
    if (rc) return rc;
    n_apply++;
    k_sub<<<W.nblocks, RED_THREADS, 0, S.st>>>(b_dev, W.w, W.t, S.n2);
    ctx->launches++;
    rc = S.norm2(W.t, W.nrm);
    if (rc) return rc;
    rc = S.fetch(W.nrm, host, 1);
    if (rc) return rc;

    res->reason = reason;
    res->iterations = it;
    res->res_norm = resn;
    res->res0_norm = res0;
    res->b_norm = bnorm;
    res->true_res_norm = std::sqrt(host[0]);
    res->n_apply = n_apply;
    {
        std::vector<double> Hs((size_t)(last_cols + 1) * std::max(last_cols, 1), 0.0);
        for (int j = 0; j < last_cols; ++j)
            for (int i = 0; i <= last_cols; ++i)
                Hs[(size_t)j * (last_cols + 1) + i] = Hraw[(size_t)j * (m + 1) + i];
        extreme_singular_values(Hs, last_cols + 1, last_cols, &res->sigma_max, &res->sigma_min);
    }
    return RTB200_OK;
}
