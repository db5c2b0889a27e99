// Node-wise block-Jacobi preconditioner for columns with an ARBITRARY angular mesh
// (hp-refined in angle: no block-circulant structure), z = (M - S)^{-1} r with the
// reference's own "preconditioner matrix" (preconditioner_matrix.py:22-267), which is
// block diagonal over (column, spatial node) with blocks
//
//     B_n = mk_n diag(wt) - ms_n G,        G = diag(wt) Phi diag(wt)   [nt x nt].
//
// For symmetric scattering kernels (any Phi(|theta - theta'|); checked on the host) all
// blocks of a signature are diagonalised by ONE orthogonal matrix: with D = diag(wt) and
// D^{-1/2} G D^{-1/2} = Q diag(lambda) Q^T,
//
//     B_n^{-1} = W diag(1 / (mk_n - ms_n lambda_k)) W^T,      W = D^{-1/2} Q,
//
// so the batched LU of the per-node blocks (north star, subsystem 4) reduces to one
// symmetric eigen-decomposition per angular signature on the host
// (rt/tables.py:pc_eig_tables) and two dense products per node on the device -- no
// per-node factors in memory (they would take nt doubles per DoF).
//
// Table of one signature at sig_eig_off[s] (doubles): [lambda: nt][W: nt x nt, row theta]
// [W^T: nt x nt, row k].  Extinction floor as in k_pc_circ: kappa + shift / min(dx, dy).
#pragma once

constexpr int PCE_THREADS = 256;

__global__ void __launch_bounds__(PCE_THREADS)
k_pc_eig(DevMesh m, const int32_t* __restrict__ cols, const int64_t* __restrict__ sig_eig_off,
         const double* __restrict__ eig_pool, const double shift, const double* __restrict__ x,
         double* __restrict__ y) {
    extern __shared__ __align__(16) double pce_s[];
    const int c = cols[blockIdx.x];
    const int tid = threadIdx.x;
    const int nx = m.col_nx[c], ny = m.col_ny[c], nn = nx * ny;
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg], ntp = m.sig_ntp[sg];
    const int64_t off = m.col_slab_off[c];
    const double* __restrict__ lam = eig_pool + sig_eig_off[sg];
    const double* __restrict__ W = lam + nt;               // [theta][k]
    const double* __restrict__ Wt = W + (int64_t)nt * nt;  // [k][theta]
    double* const R = pce_s;                                // [nn][nt]  residual rows
    double* const Tm = pce_s + nn * nt;                     // [nn][nt]  modal coefficients
    double* const mk = Tm + nn * nt;                        // [nn] floored extinction samples
    double* const ms = mk + nn;                             // [nn]

    for (int e = tid; e < nn * nt; e += PCE_THREADS) {
        const int n = e / nt, th = e - n * nt;
        R[e] = x[off + (int64_t)n * ntp + th];
    }
    {
        const double dx = m.col_dx[c], dy = m.col_dy[c];
        const double* wx = c_w + nx * RT_MAXN;
        const double* wy = c_w + ny * RT_MAXN;
        const double* __restrict__ coef = m.coef + m.col_coef_off[c];
        for (int n = tid; n < nn; n += PCE_THREADS) {
            const int i = n / ny, j = n - i * ny;
            mk[n] = coef[n] + shift * 0.25 * dx * dy * wx[i] * wy[j] / fmin(dx, dy);
            ms[n] = coef[nn + n];
        }
    }
    __syncthreads();
    // modal coefficients: t[n][k] = (sum_theta r[n][theta] W[theta][k]) / (mk_n - ms_n lambda_k)
    for (int e = tid; e < nn * nt; e += PCE_THREADS) {
        const int n = e / nt, k = e - n * nt;
        const double* __restrict__ r = R + n * nt;
        double s = 0.0;
        for (int th = 0; th < nt; ++th) s = fma(r[th], W[(int64_t)th * nt + k], s);
        Tm[e] = s / (mk[n] - ms[n] * lam[k]);
    }
    __syncthreads();
    // back: z[n][theta] = sum_k t[n][k] W[theta][k]
    for (int e = tid; e < nn * ntp; e += PCE_THREADS) {
        const int n = e / ntp, th = e - n * ntp;
        double s = 0.0;
        if (th < nt) {
            const double* __restrict__ t = Tm + n * nt;
            for (int k = 0; k < nt; ++k) s = fma(t[k], Wt[(int64_t)k * nt + th], s);
        }
        y[off + (int64_t)n * ntp + th] = s;
    }
}

static int launch_pc_eig(rtb200_ctx* ctx, const Bucket& b, const double* x, double* y, cudaStream_t st) {
    if (!ctx->pc_eig_pool)
        return rt_fail(ctx, RTB200_ESTATE,
                       "preconditioner on a non-uniform angular mesh needs the eigen tables "
                       "(rtb200_pc_set_eig)");
    const size_t smem = sizeof(double) * ((size_t)2 * b.pc_max_rows_nt + 2 * b.pc_max_nn);
    if (smem > 200 * 1024)
        return rt_fail(ctx, RTB200_EUNSUPPORTED, "preconditioner: column slab too large for shared memory");
    cudaError_t e = cudaFuncSetAttribute(k_pc_eig, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_pc_eig, smem)");
    k_pc_eig<<<b.count, PCE_THREADS, smem, st>>>(ctx->dm, b.cols_dev, ctx->pc_eig_off, ctx->pc_eig_pool,
                                                 ctx->pc_shift, x, y);
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return rt_cuda(ctx, e, "k_pc_eig launch");
    return RTB200_OK;
}
