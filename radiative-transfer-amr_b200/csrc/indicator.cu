// AMR error-indicator reductions on the device-resident solution (sm_100a).
//
// After a solve the reference pulls the solution into a Projection and walks it
// with Python loops to decide where to refine:
//
//   error_angular_jump   (src/amr/Error_Indicator/error/error_angular_jump.py:14-138)
//       top / bottom trace of every angular cell (Lagrange basis at +1 / -1 of the
//       pulled-back cell), squared jumps against the lower / upper neighbour cell,
//       integrated over the column's LGL nodes:
//           err2[c][K] = 1/4 sum_ij w_i w_j 1/2 ((bot_K - top_lo)^2 + (top_K - bot_up)^2)
//   cell_intg_xy         (src/dg/projection/cell_intg_xy.py:14-41)   input of cell_hp_steer
//           I_xy[c][theta] = (dx dy / 4) sum_ij w_i w_j u[i][j][theta]
//   col_intg_th          (src/dg/projection/col_intg_th.py:14-43)    input of col_hp_steer
//           I_th[c][node]  = sum_theta (dtheta/2) w_a u[node][theta]
//
// k_err_angular does the three reductions in one launch on the slab-order vector
// the solver leaves on the device (one block per column; the slab is read from
// HBM once, 8 B/DoF, the re-reads of phases 2-3 hit L1/L2).  What is left for the
// host are the thresholds and the Legendre-coefficient test of the flagged cells
// (rt_b200.amr.Error_Indicator).
#include "rtb200_internal.cuh"

namespace {

constexpr int EI_THREADS = 128;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(EI_THREADS)
k_err_angular(DevMesh m, const double* __restrict__ u, const int32_t* __restrict__ cell_lo,
              const int32_t* __restrict__ cell_up, const double* __restrict__ lag_bot,
              const double* __restrict__ lag_top, const double* __restrict__ lgl_w,
              const int64_t* __restrict__ col_cell_off, const int64_t* __restrict__ col_th_off,
              double* __restrict__ cell_err2, double* __restrict__ intg_xy,
              double* __restrict__ intg_th, const int trace_doubles, const int stage) {
    extern __shared__ __align__(16) double ei_s[];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = EI_THREADS / 32;
    const int nx = m.col_nx[c], ny = m.col_ny[c], nn = nx * ny;
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg], ntp = m.sig_ntp[sg];
    const int64_t tp0 = m.sig_th0[sg];
    const int cell0 = m.sig_cell0[sg], ncell = m.sig_ncell[sg];
    const double* uc = u + m.col_slab_off[c];
    const double* __restrict__ wx = lgl_w + nx * RT_MAXN;
    const double* __restrict__ wy = lgl_w + ny * RT_MAXN;
    double* const tops = ei_s;                       // [ncell][nn]
    double* const bots = ei_s + (size_t)ncell * nn;  // [ncell][nn]
    if (stage) {
        // the column's slab, read from HBM once with coalesced loads; the three reductions
        // below then run from shared memory (host: every column fits, else stage = 0)
        // (16-byte asynchronous copies, all in flight at once: one memory latency per column;
        // slab offsets and row strides are even, so every chunk is 16-byte aligned)
        double* const us = ei_s + 2 * (size_t)trace_doubles;
        const int chunks = (nn * ntp) >> 1;
        for (int e = tid; e < chunks; e += EI_THREADS) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(us + 2 * e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(uc + 2 * e));
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        uc = us;
    }

    // ---- traces at the upper / lower end of every cell (error_angular_jump.py:31-51) ---- //
    // (cell index fastest: neighbouring threads read neighbouring chunks of one row)
    for (int e = tid; e < ncell * nn; e += EI_THREADS) {
        const int n = e / ncell, K = e - n * ncell;
        const int nth = m.cell_nth[cell0 + K], tho = m.cell_thoff[cell0 + K];
        const double* __restrict__ lb = lag_bot + nth * RTB200_MAX_NTH;
        const double* __restrict__ lt = lag_top + nth * RTB200_MAX_NTH;
        const double* row = uc + (int64_t)n * ntp + tho;
        double t = 0.0, b = 0.0;
        for (int a = 0; a < nth; ++a) {
            const double v = row[a];
            t = fma(v, lt[a], t);
            b = fma(v, lb[a], b);
        }
        tops[K * nn + n] = t;
        bots[K * nn + n] = b;
    }
    __syncthreads();

    // ---- squared jumps, integrated over the spatial nodes (:56-96): warp per cell -------- //
    for (int K = warp; K < ncell; K += NW) {
        const int lo = cell_lo[cell0 + K], up = cell_up[cell0 + K];
        double s = 0.0;
        for (int n = lane; n < nn; n += 32) {
            const int i = n / ny, j = n - i * ny;
            const double dl = bots[K * nn + n] - tops[lo * nn + n];
            const double du = tops[K * nn + n] - bots[up * nn + n];
            s = fma(wx[i] * wy[j], 0.5 * (dl * dl + du * du), s);
        }
        s = warp_sum(s);
        if (lane == 0) cell_err2[col_cell_off[c] + K] = 0.25 * s;
    }

    // ---- spatial integral per angle (cell_intg_xy.py:24-39): thread per theta ------------ //
    const double dcoef = 0.25 * m.col_dx[c] * m.col_dy[c];
    for (int th = tid; th < nt; th += EI_THREADS) {
        double s = 0.0;
        for (int n = 0; n < nn; ++n) {
            const int i = n / ny, j = n - i * ny;
            s = fma(wx[i] * wy[j], uc[(int64_t)n * ntp + th], s);
        }
        intg_xy[col_th_off[c] + th] = dcoef * s;
    }

    // ---- angular integral per node (col_intg_th.py:22-41): warp per node ----------------- //
    const int64_t n0 = m.col_coef_off[c] >> 1;
    for (int n = warp; n < nn; n += NW) {
        const double* row = uc + (int64_t)n * ntp;
        double s = 0.0;
        for (int th = lane; th < nt; th += 32) s = fma(m.th_wt[tp0 + th], row[th], s);
        s = warp_sum(s);
        if (lane == 0) intg_th[n0 + n] = s;
    }
}

// ---- spatial jump (error_spatial_jump.py:30-252) ------------------------------------ //
// One block per column.  Work item = (face F, neighbour slot, own cell K0, overlapping
// neighbour cell k): a warp takes an item, its lanes the points (p, q) of the common
// quadrature [nB x nH]; both traces are interpolated there with the pooled Lagrange
// tables (amr/spatial_jump.py) and the weighted squared difference is accumulated.
//   col_err2[c] = 1 / (perimeter 2 pi) sum_items (dA dG / 4) sum_pq wB_p wH_q (u0 - u1)^2
__global__ void __launch_bounds__(EI_THREADS)
k_err_spatial(DevMesh m, const double* __restrict__ u, const int32_t* __restrict__ sp_off,
              const int32_t* __restrict__ ap_off, const int32_t* __restrict__ nb_tab,
              const double* __restrict__ pool, const int32_t* __restrict__ apair,
              double* __restrict__ col_err2) {
    __shared__ double s_part[EI_THREADS / 32];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = EI_THREADS / 32;
    const int nx = m.col_nx[c], ny = m.col_ny[c];
    const int sg = m.col_sig[c];
    const int ntp = m.sig_ntp[sg];
    const int cell0 = m.sig_cell0[sg], ncell = m.sig_ncell[sg];
    const double* __restrict__ uc = u + m.col_slab_off[c];

    double acc = 0.0;
    const int n_items = 8 * ncell * 2;
    for (int item = warp; item < n_items; item += NW) {
        const int k = item & 1, K0 = (item >> 1) % ncell, fs = (item >> 1) / ncell;   // fs = F*2 + slot
        const int F = fs >> 1;
        const int fidx = c * 8 + fs;
        const int so = sp_off[fidx];
        if (so < 0) continue;                               // warp-uniform
        const int64_t base1 = m.face_base[fidx];
        if (base1 < 0) continue;                            // (ghost traces: single rank only)
        const int32_t* __restrict__ ap = apair + ap_off[fidx] + (K0 * 2 + k) * 4;
        const int ao = ap[0];
        if (ao < 0) continue;
        const int tho1 = ap[1], nth1 = ap[2];
        const int nth0 = m.cell_nth[cell0 + K0], tho0 = m.cell_thoff[cell0 + K0];
        const int n0 = (F & 1) ? nx : ny, n1 = m.face_n1[fidx];
        // a negative count marks identity tables (same nodes on both sides)
        const int nBs = nb_tab[fidx], nHs = ap[3];
        const int nB = nBs < 0 ? -nBs : nBs, nH = nHs < 0 ? -nHs : nHs;
        const bool conforming = nBs < 0 && nHs < 0;
        const int rs1 = m.face_rstride[fidx];
        // own face rows: F = 0: i = nx-1; 1: j = ny-1; 2: i = 0; 3: j = 0
        const int node0 = F == 0 ? (nx - 1) * ny : (F == 1 ? ny - 1 : 0);
        const int step0 = (F & 1) ? ny : 1;
        const double* __restrict__ S0 = pool + so;
        const double* __restrict__ S1 = S0 + n0 * nB;
        const double* __restrict__ wB = S1 + n1 * nB;
        const double dA = wB[nB];
        const double* __restrict__ T0 = pool + ao;
        const double* __restrict__ T1 = T0 + nth0 * nH;
        const double* __restrict__ wH = T1 + nth1 * nH;
        const double dG = wH[nH];
        const double* __restrict__ u0 = uc + (int64_t)node0 * ntp + tho0;
        const double* __restrict__ u1 = u + base1 + tho1;
        double s = 0.0;
        if (conforming) {           // warp-uniform: the jump is the nodal difference
            for (int e = lane; e < nB * nH; e += 32) {
                const int p = e / nH, q = e - p * nH;
                const double d = u0[(int64_t)p * step0 * ntp + q] - u1[(int64_t)p * rs1 + q];
                s = fma(wB[p] * wH[q], d * d, s);
            }
            acc = fma(0.25 * dA * dG, s, acc);
            continue;
        }
        for (int e = lane; e < nB * nH; e += 32) {
            const int p = e / nH, q = e - p * nH;
            double d = 0.0;
            for (int j = 0; j < n0; ++j) {
                const double* __restrict__ row = u0 + (int64_t)j * step0 * ntp;
                double t = 0.0;
                for (int a = 0; a < nth0; ++a) t = fma(T0[a * nH + q], row[a], t);
                d = fma(S0[j * nB + p], t, d);
            }
            for (int j = 0; j < n1; ++j) {
                const double* __restrict__ row = u1 + (int64_t)j * rs1;
                double t = 0.0;
                for (int a = 0; a < nth1; ++a) t = fma(T1[a * nH + q], row[a], t);
                d = fma(-S1[j * nB + p], t, d);
            }
            s = fma(wB[p] * wH[q], d * d, s);
        }
        acc = fma(0.25 * dA * dG, s, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < NW; ++w) tot += s_part[w];
        const double perim = 2.0 * m.col_dx[c] + 2.0 * m.col_dy[c];
        col_err2[c] = tot / (perim * 2.0 * 3.14159265358979323846);
    }
}

// ---- high-resolution error (error_high_resolution.py:83-216) --------------------------- //
// The low-resolution solution interpolated to the nodes of a p-refined copy of the same
// mesh (same columns and cells, more nodes: ref_mesh(form = "p")), compared there with the
// high-resolution solution:
//   cell_err2[c][K] = (dx dy dth / 8) sum_pqr w_p w_q w_r (u_hr - sum_ija phi_ip psi_jq xsi_ar u_ija)^2
//   col_hr2[c]      = (dx dy / 4) sum_K (dth / 2) sum_pqr w_p w_q w_r u_hr^2     (the norm, :83-113)
// One block per column, the cells one after the other: the low-resolution cell is staged in
// shared memory, a thread owns high-resolution nodes (p, q, r); block reduction in a fixed
// order.  mlo / mhi: the two meshes' device tables (columns and cells correspond one to one).
__global__ void __launch_bounds__(EI_THREADS)
k_err_hires(DevMesh mlo, DevMesh mhi, const double* __restrict__ ulo, const double* __restrict__ uhi,
            const int32_t* __restrict__ lgl_ip_off, const int32_t* __restrict__ lg_ip_off,
            const double* __restrict__ ip_pool, const double* __restrict__ lgl_w,
            const double* __restrict__ lg_w, const double* __restrict__ cell_dth,
            const int64_t* __restrict__ col_cell_off, double* __restrict__ cell_err2,
            double* __restrict__ col_hr2) {
    extern __shared__ __align__(16) double hr_s[];
    __shared__ double s_red[2][EI_THREADS / 32];
    const int c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = EI_THREADS / 32;
    const int nx = mlo.col_nx[c], ny = mlo.col_ny[c];
    const int px = mhi.col_nx[c], py = mhi.col_ny[c];
    const int sgl = mlo.col_sig[c], sgh = mhi.col_sig[c];
    const int ntpl = mlo.sig_ntp[sgl], ntph = mhi.sig_ntp[sgh];
    const int cl0 = mlo.sig_cell0[sgl], ch0 = mhi.sig_cell0[sgh], ncell = mlo.sig_ncell[sgl];
    const double* __restrict__ ul = ulo + mlo.col_slab_off[c];
    const double* __restrict__ uh = uhi + mhi.col_slab_off[c];
    const double* __restrict__ phi = ip_pool + lgl_ip_off[nx * (RT_MAXN + 1) + px];   // [nx][px]
    const double* __restrict__ psi = ip_pool + lgl_ip_off[ny * (RT_MAXN + 1) + py];   // [ny][py]
    const double* __restrict__ wx = lgl_w + px * RT_MAXN;
    const double* __restrict__ wy = lgl_w + py * RT_MAXN;
    const double dxy = mlo.col_dx[c] * mlo.col_dy[c];
    double hr2_col = 0.0;
    for (int K = 0; K < ncell; ++K) {
        const int nth = mlo.cell_nth[cl0 + K], thol = mlo.cell_thoff[cl0 + K];
        const int pth = mhi.cell_nth[ch0 + K], thoh = mhi.cell_thoff[ch0 + K];
        const double* __restrict__ xsi = ip_pool + lg_ip_off[nth * (RTB200_MAX_NTH + 1) + pth];   // [nth][pth]
        const double* __restrict__ wth = lg_w + pth * RTB200_MAX_NTH;
        const double dth = cell_dth[cl0 + K];
        __syncthreads();                                   // previous cell's reads of hr_s are done
        for (int e = tid; e < nx * ny * nth; e += EI_THREADS) {
            const int n = e / nth, a = e - n * nth;
            hr_s[e] = ul[(int64_t)n * ntpl + thol + a];
        }
        __syncthreads();
        double s_err = 0.0, s_hr2 = 0.0;
        for (int e = tid; e < px * py * pth; e += EI_THREADS) {
            const int pq = e / pth, r = e - pq * pth;
            const int p = pq / py, q = pq - p * py;
            double v = 0.0;
            for (int i = 0; i < nx; ++i) {
                double vi = 0.0;
                for (int j = 0; j < ny; ++j) {
                    const double* __restrict__ row = hr_s + (i * ny + j) * nth;
                    double vj = 0.0;
                    for (int a = 0; a < nth; ++a) vj = fma(xsi[a * pth + r], row[a], vj);
                    vi = fma(psi[j * py + q], vj, vi);
                }
                v = fma(phi[i * px + p], vi, v);
            }
            const double h = uh[(int64_t)pq * ntph + thoh + r];
            const double w = wx[p] * wy[q] * wth[r];
            s_err = fma(w, (h - v) * (h - v), s_err);
            s_hr2 = fma(w, h * h, s_hr2);
        }
        s_err = warp_sum(s_err);
        s_hr2 = warp_sum(s_hr2);
        if (lane == 0) {
            s_red[0][warp] = s_err;
            s_red[1][warp] = s_hr2;
        }
        __syncthreads();
        if (tid == 0) {
            double te = 0.0, th2 = 0.0;
            for (int w = 0; w < NW; ++w) {
                te += s_red[0][w];
                th2 += s_red[1][w];
            }
            cell_err2[col_cell_off[c] + K] = 0.125 * dxy * dth * te;
            hr2_col += 0.25 * dxy * (0.5 * dth) * th2;
        }
    }
    if (tid == 0) col_hr2[c] = hr2_col;
}

}  // namespace

extern "C" int rtb200_err_high_resolution(rtb200_ctx* ctx_lo, const double* u_lo_slab_dev,
                                          rtb200_ctx* ctx_hi, const double* u_hi_slab_dev,
                                          const int32_t* lgl_ip_off_dev, const int32_t* lg_ip_off_dev,
                                          const double* ip_pool_dev, const double* lgl_w_dev,
                                          const double* lg_w_dev, const double* cell_dth_dev,
                                          const int64_t* col_cell_off_dev, int32_t max_cell_dofs,
                                          double* cell_err2_dev, double* col_hr2_dev, void* stream) {
    if (!ctx_lo || !ctx_hi) return RTB200_EINVAL;
    rtb200_ctx* ctx = ctx_lo;
    if (!ctx_lo->have_mesh || !ctx_hi->have_mesh)
        return rt_fail(ctx, RTB200_ESTATE, "rtb200_err_high_resolution: no mesh set");
    if (!u_lo_slab_dev || !u_hi_slab_dev || !lgl_ip_off_dev || !lg_ip_off_dev || !ip_pool_dev ||
        !lgl_w_dev || !lg_w_dev || !cell_dth_dev || !col_cell_off_dev || !cell_err2_dev || !col_hr2_dev)
        return rt_fail(ctx, RTB200_EINVAL, "rtb200_err_high_resolution: null pointer");
    if (ctx_lo->dm.n_col != ctx_hi->dm.n_col || ctx_lo->device != ctx_hi->device)
        return rt_fail(ctx, RTB200_EINVAL,
                       "rtb200_err_high_resolution: the two meshes must have the same columns, on one device");
    if (ctx_lo->nranks > 1 || ctx_hi->nranks > 1)
        return rt_fail(ctx, RTB200_EUNSUPPORTED, "rtb200_err_high_resolution: single rank only");
    if (max_cell_dofs <= 0 || (size_t)max_cell_dofs * sizeof(double) > 200 * 1024)
        return rt_fail(ctx, RTB200_EUNSUPPORTED, "rtb200_err_high_resolution: cell too large for shared memory");
    if (ctx->dm.n_col == 0) return RTB200_OK;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t smem = sizeof(double) * (size_t)max_cell_dofs;
    if (smem > 48 * 1024)
        RT_CUDA(ctx, cudaFuncSetAttribute(k_err_hires, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_err_hires<<<ctx->dm.n_col, EI_THREADS, smem, (cudaStream_t)stream>>>(
        ctx_lo->dm, ctx_hi->dm, u_lo_slab_dev, u_hi_slab_dev, lgl_ip_off_dev, lg_ip_off_dev, ip_pool_dev,
        lgl_w_dev, lg_w_dev, cell_dth_dev, col_cell_off_dev, cell_err2_dev, col_hr2_dev);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}

extern "C" int rtb200_err_spatial_jump(rtb200_ctx* ctx, const double* u_slab_dev,
                                       const int32_t* sp_off_dev, const int32_t* ap_off_dev,
                                       const int32_t* nb_dev, const double* pool_dev,
                                       const int32_t* apair_dev, double* col_err2_dev,
                                       void* stream) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "rtb200_err_spatial_jump: no mesh set");
    if (!u_slab_dev || !sp_off_dev || !ap_off_dev || !nb_dev || !pool_dev || !apair_dev || !col_err2_dev)
        return rt_fail(ctx, RTB200_EINVAL, "rtb200_err_spatial_jump: null pointer");
    if (ctx->nranks > 1)
        return rt_fail(ctx, RTB200_EUNSUPPORTED, "rtb200_err_spatial_jump: single rank only");
    if (ctx->dm.n_col == 0) return RTB200_OK;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    k_err_spatial<<<ctx->dm.n_col, EI_THREADS, 0, (cudaStream_t)stream>>>(
        ctx->dm, u_slab_dev, sp_off_dev, ap_off_dev, nb_dev, pool_dev, apair_dev, col_err2_dev);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}

extern "C" int rtb200_err_angular_jump(rtb200_ctx* ctx, const double* u_slab_dev,
                                       const int32_t* cell_lo_dev, const int32_t* cell_up_dev,
                                       const double* lag_bot_dev, const double* lag_top_dev,
                                       const double* lgl_w_dev, const int64_t* col_cell_off_dev,
                                       const int64_t* col_th_off_dev, int32_t max_cell_nodes,
                                       int32_t max_slab_doubles, double* cell_err2_dev,
                                       double* cell_intg_xy_dev, double* col_intg_th_dev,
                                       void* stream) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "rtb200_err_angular_jump: no mesh set");
    if (!u_slab_dev || !cell_lo_dev || !cell_up_dev || !lag_bot_dev || !lag_top_dev || !lgl_w_dev ||
        !col_cell_off_dev || !col_th_off_dev || !cell_err2_dev || !cell_intg_xy_dev || !col_intg_th_dev)
        return rt_fail(ctx, RTB200_EINVAL, "rtb200_err_angular_jump: null pointer");
    if (max_cell_nodes <= 0) return rt_fail(ctx, RTB200_EINVAL, "rtb200_err_angular_jump: max_cell_nodes <= 0");
    size_t smem = 2 * sizeof(double) * (size_t)max_cell_nodes;
    if (smem > 200 * 1024)
        return rt_fail(ctx, RTB200_EUNSUPPORTED,
                       "rtb200_err_angular_jump: a column with more than 12800 (cell, node) pairs");
    // stage the slab in shared memory when every column fits next to its traces and at
    // least two blocks stay resident per SM
    const int stage = max_slab_doubles > 0 &&
                      smem + sizeof(double) * (size_t)max_slab_doubles <= 100 * 1024;
    if (stage) smem += sizeof(double) * (size_t)max_slab_doubles;
    if (ctx->dm.n_col == 0) return RTB200_OK;
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (smem > 48 * 1024)
        RT_CUDA(ctx, cudaFuncSetAttribute(k_err_angular, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_err_angular<<<ctx->dm.n_col, EI_THREADS, smem, (cudaStream_t)stream>>>(
        ctx->dm, u_slab_dev, cell_lo_dev, cell_up_dev, lag_bot_dev, lag_top_dev, lgl_w_dev,
        col_cell_off_dev, col_th_off_dev, cell_err2_dev, cell_intg_xy_dev, col_intg_th_dev,
        (int)max_cell_nodes, stage);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}
