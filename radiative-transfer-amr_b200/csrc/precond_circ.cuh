// Node-wise block-Jacobi preconditioner  z = mask . (M - S)^{-1} r  for columns with
// a uniform angular mesh (block-circulant scattering, NC >= 8).
//
// M - S (the reference's "preconditioner matrix", preconditioner_matrix.py:116-267) is
// block diagonal over (column, spatial node); the block of a node is
//     B = mk diag(wt) - ms G                       (NT x NT, NT = NC P angular DoFs)
// with G block circulant in the cell index.  The FFT over the cell index that the
// apply kernel uses for G x block-diagonalises B as well: per frequency k a P x P
// complex system
//     (mk diag(wt_a) - ms Chat[k]) Z[k] = X[k],
// solved here by Gaussian elimination without pivoting (B is symmetric positive
// definite for sigma <= kappa, and so is every frequency block).  Chat[k] is recovered
// from the apply kernel's table T (rt/tables.py:circulant_spectrum):
//     Chat[0] = nc T[0].re, Chat[nc/2] = nc T[0].im, Chat[k] = 2 nc T[k].
// With the transform conventions of fft_real.cuh (forward slot k = 2 X[k] for k > 0,
// inverse returns nc x):  z = inverse( slot0: Z, slot k: Z[k] from rhs slot / 2 ) / nc.
//
// Vacuum regions (kappa = 0: M - S vanishes there, pure advection) would make the blocks
// singular, so the extinction is floored: kappa_eff = kappa + shift / min(dx, dy) -- the
// magnitude of the advection operator on the column -- with ctx->pc_shift = 1 by default
// (rtb200_pc_set_shift; 0 gives the exact (M - S)^{-1}).  Any SPD choice is a valid
// preconditioner; it only changes the iteration count.
//
// Rows of inflow-boundary DoFs are zero in r (GMRES vectors) and are zeroed again in z
// (k_mask): on boundary nodes the result is the interior part of the unrestricted
// inverse, a fixed linear approximation of (M - S)_ii^{-1} -- still a valid
// preconditioner.  One block per column; transforms shared with k_apply_circ7.
#pragma once

// ---- shuffle form of the split transforms (full spectrum in the slot layout [k][a]) ---------- //
// k_apply_circ7 moved the last radix-2 stage into its block product (E/O form, apply_circ7.cuh);
// this kernel solves one P x P system per slot and keeps the complete transforms: the two lanes
// (l, l ^ 16) of a sequence exchange NC/4 doubles with shuffles.
// ---- split real FFT: forward ------------------------------------------------ //
// xs[m] = x[2 m + t], m < NC/2.  Output: the NC/4 spectrum slots this lane owns, in
// the slot convention of fft_real.cuh for length NC (slot 0 = (X[0], X[NC/2]),
// slot k = 2 X[k]):
//   pair p (0 .. NC/8): k = 2 p + t ->  oa[p] = slot k,  ob[p] = slot NC/2 - k,
//   except (t, p) = (0, 0):             oa[0] = slot 0,  ob[0] = slot NC/4.
// With E / O the half-length spectra of the even / odd samples and w = exp(-2 pi i / NC):
//   X[k] = E[k] + w^k O[k],  X[NC/2 - k] = conj(E[k] - w^k O[k]).
template <int NC>
__device__ __forceinline__ void split_forward(const double (&xs)[NC / 2], const int t,
                                              double2 (&oa)[NC / 8], double2 (&ob)[NC / 8]) {
    constexpr int Q = NC / 4, NP = NC / 8;
    double hr[Q], hi[Q];
    rtfft::Real<NC / 2>::forward(xs, hr, hi);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        // lane t = 0 (even samples, E) sends E[2p+1]; lane t = 1 (O) sends O[2p]
        const double sr = t ? hr[2 * p] : hr[2 * p + 1];
        const double si = t ? hi[2 * p] : hi[2 * p + 1];
        const double rr = __shfl_xor_sync(0xffffffffu, sr, 16);
        const double ri = __shfl_xor_sync(0xffffffffu, si, 16);
        // generic combination at k = 2p + t
        const double er = t ? rr : hr[2 * p], ei = t ? ri : hi[2 * p];
        const double orr = t ? hr[2 * p + 1] : rr, oi = t ? hi[2 * p + 1] : ri;
        const double c = t ? c_fft_cos[(2 * p + 1) * (64 / NC)] : c_fft_cos[(2 * p) * (64 / NC)];
        const double s = t ? c_fft_sin[(2 * p + 1) * (64 / NC)] : c_fft_sin[(2 * p) * (64 / NC)];
        const double wr = fma(c, orr, s * oi), wi = fma(c, oi, -(s * orr));
        double2 a = make_double2(er + wr, ei + wi);
        double2 b = make_double2(er - wr, wi - ei);
        if (p == 0) {
            // t = 0: E = (E[0], E[Q]) and O = (O[0], O[Q]) are the packed real bins
            const double2 sa = make_double2(hr[0] + rr, hr[0] - rr);
            const double2 sb = make_double2(2.0 * hi[0], -2.0 * ri);
            a = t ? a : sa;
            b = t ? b : sb;
        }
        oa[p] = a;
        ob[p] = b;
    }
}

// ---- split real FFT: inverse ------------------------------------------------ //
// ia / ib: the slots this lane owns (same ownership as split_forward), holding the
// TRUE spectrum Y.  Output xs[m] = NC * x[2 m + t].
//   A[k] = Y[k] + conj(Y[NC/2 - k])           -> even samples
//   B[k] = (Y[k] - conj(Y[NC/2 - k])) w^{-k}  -> odd samples
template <int NC>
__device__ __forceinline__ void split_inverse(const double2 (&ia)[NC / 8], const double2 (&ib)[NC / 8],
                                              const int t, double (&xs)[NC / 2]) {
    constexpr int Q = NC / 4, NP = NC / 8;
    double hr[Q], hi[Q];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const double2 yk = ia[p], ym = ib[p];
        double ar = yk.x + ym.x, ai = yk.y - ym.y;
        const double dr = yk.x - ym.x, di = yk.y + ym.y;
        const double c = t ? c_fft_cos[(2 * p + 1) * (64 / NC)] : c_fft_cos[(2 * p) * (64 / NC)];
        const double s = t ? c_fft_sin[(2 * p + 1) * (64 / NC)] : c_fft_sin[(2 * p) * (64 / NC)];
        double br = fma(dr, c, -(di * s)), bi = fma(dr, s, di * c);
        if (p == 0) {
            // t = 0: yk = (Y[0], Y[NC/2]), ym = Y[NC/4]
            const double a0 = yk.x + yk.y, aq = 2.0 * ym.x;
            const double b0 = yk.x - yk.y, bq = -2.0 * ym.y;
            ar = t ? ar : a0;
            ai = t ? ai : aq;
            br = t ? br : b0;
            bi = t ? bi : bq;
        }
        // lane 0 keeps A and sends B[2p]; lane 1 keeps B and sends A[2p+1]
        const double sr = t ? ar : br, si = t ? ai : bi;
        const double rr = __shfl_xor_sync(0xffffffffu, sr, 16);
        const double ri = __shfl_xor_sync(0xffffffffu, si, 16);
        hr[2 * p] = t ? rr : ar;
        hi[2 * p] = t ? ri : ai;
        hr[2 * p + 1] = t ? br : rr;
        hi[2 * p + 1] = t ? bi : ri;
    }
    rtfft::Real<NC / 2>::inverse(hr, hi, xs);
}

template <typename Cfg>
__device__ __noinline__ void fft_f7(const double* __restrict__ U, double* __restrict__ W,
                                    const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, NP = NC / 8, SU = Cfg::SU, SW = Cfg::SW;
    constexpr int NN = Cfg::NN;
    if (warp >= Cfg::WORK_WARPS) return;                        // warp-uniform
    const int n0 = (warp * NN) / Cfg::WORK_WARPS;
    const int nw = ((warp + 1) * NN) / Cfg::WORK_WARPS - n0;
    if (nw == 0) return;
    const int t = lane >> 4, sq = lane & 15;
    const bool seq_on = sq < nw * P;
    const int dn = seq_on ? sq / P : 0, a = seq_on ? sq - dn * P : 0;
    const int node = n0 + dn;
    double xs[H];
    const double* src = U + node * SU + a + t * P;
#pragma unroll
    for (int mm = 0; mm < H; ++mm) xs[mm] = src[2 * mm * P];
    double2 oa[NP], ob[NP];
    split_forward<NC>(xs, t, oa, ob);
    if (seq_on) {
        double2* const wa = reinterpret_cast<double2*>(W + node * SW) + a + t * P;   // slot 2p + t
        double2* const wb = reinterpret_cast<double2*>(W + node * SW) + a - t * P;   // slot NC/2 - 2p - t
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            wa[2 * p * P] = oa[p];
            const int sb = (p == 0 && t == 0) ? (NC / 4) : (NC / 2 - 2 * p);
            wb[sb * P] = ob[p];
        }
    }
}

template <typename Cfg>
__device__ __noinline__ void fft_i7(double* __restrict__ W, const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, NP = NC / 8, SW = Cfg::SW;
    constexpr int NN = Cfg::NN;
    if (warp >= Cfg::WORK_WARPS) return;                        // warp-uniform
    const int n0 = (warp * NN) / Cfg::WORK_WARPS;
    const int nw = ((warp + 1) * NN) / Cfg::WORK_WARPS - n0;
    if (nw == 0) return;
    const int t = lane >> 4, sq = lane & 15;
    const bool seq_on = sq < nw * P;
    const int dn = seq_on ? sq / P : 0, a = seq_on ? sq - dn * P : 0;
    const int node = n0 + dn;
    const double2* const wa = reinterpret_cast<const double2*>(W + node * SW) + a + t * P;
    const double2* const wb = reinterpret_cast<const double2*>(W + node * SW) + a - t * P;
    double2 ia[NP], ib[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        ia[p] = wa[2 * p * P];
        const int sb = (p == 0 && t == 0) ? (NC / 4) : (NC / 2 - 2 * p);
        ib[p] = wb[sb * P];
    }
    __syncwarp();
    double xs[H];
    split_inverse<NC>(ia, ib, t, xs);
    if (seq_on) {
        double* dst = W + n0 * SW + dn * Cfg::SWR + a + t * P;      // real-row layout, see Circ7Cfg::wrow
#pragma unroll
        for (int mm = 0; mm < H; ++mm) dst[2 * mm * P] = xs[mm];
    }
}



template <int P>
__device__ __forceinline__ void solve_complex(double (&ar)[P][P], double (&ai)[P][P], double (&br)[P],
                                              double (&bi)[P]) {
    double pr[P], pi[P];
#pragma unroll
    for (int c = 0; c < P; ++c) {
        const double d = 1.0 / fma(ar[c][c], ar[c][c], ai[c][c] * ai[c][c]);
        pr[c] = ar[c][c] * d;
        pi[c] = -(ai[c][c] * d);
#pragma unroll
        for (int r = c + 1; r < P; ++r) {
            const double lr = ar[r][c] * pr[c] - ai[r][c] * pi[c];
            const double li = ar[r][c] * pi[c] + ai[r][c] * pr[c];
#pragma unroll
            for (int q = c + 1; q < P; ++q) {
                ar[r][q] -= lr * ar[c][q] - li * ai[c][q];
                ai[r][q] -= lr * ai[c][q] + li * ar[c][q];
            }
            br[r] -= lr * br[c] - li * bi[c];
            bi[r] -= lr * bi[c] + li * br[c];
        }
    }
#pragma unroll
    for (int c = P - 1; c >= 0; --c) {
        double sr = br[c], si = bi[c];
#pragma unroll
        for (int q = c + 1; q < P; ++q) {
            sr -= ar[c][q] * br[q] - ai[c][q] * bi[q];
            si -= ar[c][q] * bi[q] + ai[c][q] * br[q];
        }
        br[c] = sr * pr[c] - si * pi[c];
        bi[c] = sr * pi[c] + si * pr[c];
    }
}

template <int P>
__device__ __forceinline__ void solve_real(double (&a)[P][P], double (&b)[P]) {
    double p[P];
#pragma unroll
    for (int c = 0; c < P; ++c) {
        p[c] = 1.0 / a[c][c];
#pragma unroll
        for (int r = c + 1; r < P; ++r) {
            const double l = a[r][c] * p[c];
#pragma unroll
            for (int q = c + 1; q < P; ++q) a[r][q] -= l * a[c][q];
            b[r] -= l * b[c];
        }
    }
#pragma unroll
    for (int c = P - 1; c >= 0; --c) {
        double s = b[c];
#pragma unroll
        for (int q = c + 1; q < P; ++q) s -= a[c][q] * b[q];
        b[c] = s * p[c];
    }
}

template <int N, int NC, int P>
__global__ void __launch_bounds__(Circ7Cfg<N, NC, P>::THREADS, Circ7Cfg<N, NC, P>::MINB)
k_pc_circ(DevMesh m, const int32_t* __restrict__ cols, int sg, const double shift,
          const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ circ_tab) {
    using Cfg = Circ7Cfg<N, NC, P>;
    constexpr int NN = Cfg::NN, NT = Cfg::NT, NTP = Cfg::NTP, SU = Cfg::SU, SW = Cfg::SW, M = Cfg::M;
    constexpr int THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double smem_pc[];
    double* const U = smem_pc;
    double* const W = smem_pc + Cfg::U_DOUBLES;
    __shared__ double s_coef[2 * NN];
    __shared__ double s_wt[P];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c = cols[blockIdx.x];
    const int64_t off = m.col_slab_off[c];
    for (int e = tid; e < NN * NT; e += THREADS) {
        const int node = e / NT, th = e - node * NT;
        U[node * SU + th] = x[off + (int64_t)node * NTP + th];
    }
    for (int e = tid; e < 2 * NN; e += THREADS) {
        double v = m.coef[m.col_coef_off[c] + e];
        if (e < NN) {                // mk = (dx dy / 4) w_i w_j kappa: floor kappa by shift / min(dx, dy)
            const double dx = m.col_dx[c], dy = m.col_dy[c];
            const double* wq = c_w + N * RT_MAXN;
            v += shift * 0.25 * dx * dy * wq[e / N] * wq[e % N] / fmin(dx, dy);
        }
        s_coef[e] = v;
    }
    if (tid < P) s_wt[tid] = m.th_wt[m.sig_th0[sg] + tid];
    __syncthreads();
    fft_f7<Cfg>(U, W, warp, lane);
    __syncthreads();
    {
        const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
        const bool on = lane < NN;
        const int node = on ? lane : 0;
        const double mk = s_coef[node], ms = s_coef[NN + node];
        double2* const wn = reinterpret_cast<double2*>(W) + node * (SW / 2);
        const double2* __restrict__ T = reinterpret_cast<const double2*>(circ_tab);
        if (warp_u < Cfg::WORK_WARPS) {
#pragma unroll 1
            for (int k = warp_u; k < M; k += Cfg::WORK_WARPS) {
                double2* const w2 = wn + k * P;
                const double2* __restrict__ tk = T + k * P * P;
                if (k == 0) {
                    double a0[P][P], ah[P][P], b0[P], bh[P];
#pragma unroll
                    for (int r = 0; r < P; ++r) {
                        const double2 b = w2[r];
                        b0[r] = b.x;
                        bh[r] = b.y;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            const double2 t = __ldg(&tk[r * P + q]);
                            a0[r][q] = -(ms * (double)NC) * t.x + (r == q ? mk * s_wt[r] : 0.0);
                            ah[r][q] = -(ms * (double)NC) * t.y + (r == q ? mk * s_wt[r] : 0.0);
                        }
                    }
                    solve_real<P>(a0, b0);
                    solve_real<P>(ah, bh);
                    if (on) {
#pragma unroll
                        for (int r = 0; r < P; ++r) w2[r] = make_double2(b0[r], bh[r]);
                    }
                } else {
                    double ar[P][P], ai[P][P], br[P], bi[P];
#pragma unroll
                    for (int r = 0; r < P; ++r) {
                        const double2 b = w2[r];
                        br[r] = 0.5 * b.x;
                        bi[r] = 0.5 * b.y;
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            const double2 t = __ldg(&tk[r * P + q]);
                            ar[r][q] = -(ms * (2.0 * NC)) * t.x + (r == q ? mk * s_wt[r] : 0.0);
                            ai[r][q] = -(ms * (2.0 * NC)) * t.y;
                        }
                    }
                    solve_complex<P>(ar, ai, br, bi);
                    if (on) {
#pragma unroll
                        for (int r = 0; r < P; ++r) w2[r] = make_double2(br[r], bi[r]);
                    }
                }
            }
        }
    }
    __syncthreads();
    fft_i7<Cfg>(W, warp, lane);
    __syncthreads();
    for (int e = tid; e < NN * NTP; e += THREADS) {
        const int node = e / NTP, th = e - node * NTP;
        y[off + (int64_t)node * NTP + th] = th < NT ? W[Cfg::wrow(node) + th] * (1.0 / NC) : 0.0;
    }
}

template <int N, int NC, int P>
static int launch_pc_circ(rtb200_ctx* ctx, const Bucket& b, const double* x, double* y, cudaStream_t st) {
    if constexpr (NC >= 8) {
        using Cfg = Circ7Cfg<N, NC, P>;
        constexpr size_t smem = sizeof(double) * (size_t)(Cfg::U_DOUBLES + Cfg::W_DOUBLES);
        auto kern = k_pc_circ<N, NC, P>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_pc_circ, smem)");
        kern<<<b.count, Cfg::THREADS, smem, st>>>(ctx->dm, b.cols_dev, b.circ_sig, ctx->pc_shift, x, y,
                                                  b.circ_tab_dev);
        ctx->launches++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return rt_cuda(ctx, e, "k_pc_circ launch");
        return RTB200_OK;
    } else {
        return rt_fail(ctx, RTB200_EUNSUPPORTED, "block-Jacobi preconditioner needs >= 8 angular cells");
    }
}

// buckets served by the block-circulant kernel; the others use the eigen tables
static bool pc_circ_bucket(const Bucket& b) {
    if (b.circ_nc < 8) return false;
#define X(N_, NC_, P_) if (b.n == N_ && b.circ_nc == NC_ && b.circ_p == P_) return true;
    RT_CIRC_COMBOS(X)
#undef X
    return false;
}

int rt_pc_setup(rtb200_ctx* ctx, cudaStream_t) {
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set");
    for (const Bucket& b : ctx->buckets) {
        if (pc_circ_bucket(b)) continue;
        for (int sg : b.sigs)
            if (ctx->pc_eig_off_host.empty() || ctx->pc_eig_off_host[sg] < 0)
                return rt_fail(ctx, RTB200_EUNSUPPORTED,
                               "the node-wise block-Jacobi preconditioner needs, for columns without a "
                               "uniform angular mesh (>= 8 cells, block-circulant scattering), the "
                               "eigen tables of rtb200_pc_set_eig (symmetric scattering kernels)");
    }
    ctx->pc_ready = true;
    return RTB200_OK;
}

int rt_pc_apply(rtb200_ctx* ctx, const double* x, double* y, cudaStream_t st) {
    if (!ctx->pc_ready) {
        int rc = rt_pc_setup(ctx, st);
        if (rc) return rc;
    }
    // buckets of an hp mesh run concurrently on the side streams, like the apply's
    // (rt_launch_apply): fork from / join to the caller's stream
    const size_t nb = ctx->buckets.size();
    const size_t nside = nb > 1 ? std::min(ctx->side_streams.size(), nb - 1) : 0;
    if (nside > 0) {
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        for (size_t s = 0; s < nside; ++s)
            RT_CUDA(ctx, cudaStreamWaitEvent(ctx->side_streams[s], ctx->ev_fork, 0));
    }
    int rc_all = RTB200_OK;
    for (size_t i = 0; i < nb && rc_all == RTB200_OK; ++i) {
        const Bucket& b = ctx->buckets[ctx->bucket_order[i]];
        const size_t slot = i % (nside + 1);
        cudaStream_t q = slot == 0 ? st : ctx->side_streams[slot - 1];
        int rc = RTB200_EUNSUPPORTED;
        if (pc_circ_bucket(b)) {
#define X(N_, NC_, P_)                                               \
    if (b.n == N_ && b.circ_nc == NC_ && b.circ_p == P_)             \
        rc = launch_pc_circ<N_, NC_, P_>(ctx, b, x, y, q);
            RT_CIRC_COMBOS(X)
#undef X
        } else {
            rc = launch_pc_eig(ctx, b, x, y, q);
        }
        rc_all = rc;
    }
    for (size_t s = 0; s < nside; ++s) {                     // always join, also after an error
        cudaEventRecord(ctx->ev_join[s], ctx->side_streams[s]);
        cudaStreamWaitEvent(st, ctx->ev_join[s], 0);
    }
    if (rc_all) return rc_all;
    if (ctx->dm.n_col > 0) {
        k_mask<<<ctx->dm.n_col, 256, 0, st>>>(ctx->dm, y, 1);
        ctx->launches++;
        RT_CUDA(ctx, cudaGetLastError());
    }
    return RTB200_OK;
}

// dispatch on the preconditioner kind (rtb200_gmres_params.pc / rtb200_pc_set_type)
int rt_pc_apply_type(rtb200_ctx* ctx, int pc_type, const double* x, double* y, cudaStream_t st) {
    if (pc_type == RTB200_PC_NODE) return rt_pc_apply(ctx, x, y, st);
    if (pc_type == RTB200_PC_SWEEP || pc_type == RTB200_PC_ELEMENT) return rt_sweep_apply(ctx, pc_type, x, y, st);
    return rt_fail(ctx, RTB200_EINVAL, "unknown preconditioner kind");
}

void rt_pc_destroy(rtb200_ctx* ctx) {
    ctx->pc_ready = false;
    ctx->pc_eig_pool = nullptr;        // freed with the other tables (ctx->allocs)
    ctx->pc_eig_off = nullptr;
    ctx->pc_eig_off_host.clear();
}

extern "C" int rtb200_pc_set_eig(rtb200_ctx* ctx, const int64_t* sig_eig_off, int32_t n_sig,
                                 const double* eig_pool, int64_t n_pool) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set (call rtb200_set_mesh first)");
    if (!sig_eig_off || !eig_pool || n_sig <= 0 || n_pool <= 0)
        return rt_fail(ctx, RTB200_EINVAL, "pc_set_eig: empty tables");
    for (int s = 0; s < n_sig; ++s)
        if (sig_eig_off[s] >= n_pool) return rt_fail(ctx, RTB200_EINVAL, "pc_set_eig: offset outside the pool");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    double* pool = nullptr;
    int64_t* off = nullptr;
    RT_CUDA(ctx, cudaMalloc((void**)&pool, sizeof(double) * (size_t)n_pool));
    ctx->allocs.push_back(pool);
    RT_CUDA(ctx, cudaMalloc((void**)&off, sizeof(int64_t) * (size_t)n_sig));
    ctx->allocs.push_back(off);
    RT_CUDA(ctx, cudaMemcpy(pool, eig_pool, sizeof(double) * (size_t)n_pool, cudaMemcpyHostToDevice));
    RT_CUDA(ctx, cudaMemcpy(off, sig_eig_off, sizeof(int64_t) * (size_t)n_sig, cudaMemcpyHostToDevice));
    ctx->pc_eig_pool = pool;
    ctx->pc_eig_off = off;
    ctx->pc_eig_off_host.assign(sig_eig_off, sig_eig_off + n_sig);
    ctx->pc_ready = false;
    return RTB200_OK;
}

extern "C" int rtb200_pc_set_shift(rtb200_ctx* ctx, double shift) {
    if (!ctx) return RTB200_EINVAL;
    if (!(shift >= 0.0)) return rt_fail(ctx, RTB200_EINVAL, "pc_set_shift: shift must be >= 0");
    ctx->pc_shift = shift;
    return RTB200_OK;
}

extern "C" int rtb200_pc_apply(rtb200_ctx* ctx, const double* x_slab_dev, double* y_slab_dev, void* stream) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set (call rtb200_set_mesh first)");
    if (!x_slab_dev || !y_slab_dev) return rt_fail(ctx, RTB200_EINVAL, "null device vector");
    if (x_slab_dev == y_slab_dev) return rt_fail(ctx, RTB200_EINVAL, "pc_apply: x and y must not alias");
    return rt_pc_apply_type(ctx, ctx->pc_type, x_slab_dev, y_slab_dev, (cudaStream_t)stream);
}
