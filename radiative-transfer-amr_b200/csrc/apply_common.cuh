// Device code shared by the apply kernels (dense-scattering k_apply and the
// block-circulant k_apply_circ): async copies, the DMMA wrapper, the inflow
// mask test and the streaming terms (mass, volume convection, face fluxes) of
// one (column, theta) pair.
#pragma once
#include "rtb200_internal.cuh"

// c_w / c_s (LGL weights and S[p][i] = phi_p'(xi_i) w_i per n) are defined by the
// including translation unit (apply.cu) before this header.

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ bool dof_masked(int i, int j, int nx, int ny, int bdry, int bits) {
    return ((i == nx - 1) && (bdry & 1) && (bits & 16)) || ((j == ny - 1) && (bdry & 2) && (bits & 32)) ||
           ((i == 0) && (bdry & 4) && (bits & 64)) || ((j == 0) && (bdry & 8) && (bits & 128));
}

struct GroupTab {
    const int32_t* grp_first;   // [n_groups] first entry in grp_cols
    const int32_t* grp_count;   // [n_groups]
    const int32_t* grp_cols;    // column indices
};

// Column metadata staged in shared memory by the first 9 threads of the block.
struct ColMeta {
    long long slab_off;
    long long face_base[8];
    double hdx, hdy;
    int bdry, c;
    int rs[8], exy[8], acpl[8], n1[8];
};

__device__ __forceinline__ void load_col_meta(const DevMesh& m, const int c, ColMeta& cm, const int t) {
    if (t < 8) {
        const int fidx = c * 8 + t;
        cm.face_base[t] = m.face_base[fidx];
        cm.rs[t] = m.face_rstride[fidx];
        cm.exy[t] = m.face_exy[fidx];
        cm.acpl[t] = m.face_acpl[fidx];
        cm.n1[t] = m.face_n1[fidx];
    } else if (t == 8) {
        cm.slab_off = m.col_slab_off[c];
        cm.hdx = 0.5 * m.col_dx[c];
        cm.hdy = 0.5 * m.col_dy[c];
        cm.bdry = m.col_bdry[c];
        cm.c = c;
    }
}

struct ColMeta7 {
    ColMeta cm;
    long long coef_off;
};

__device__ __forceinline__ void load_col_meta7(const DevMesh& m, const int c, ColMeta7& cm, const int t) {
    load_col_meta(m, c, cm.cm, t);
    if (t == 8) cm.coef_off = m.col_coef_off[c];
}

// Streaming terms of column c at angular index th (< nt), for every spatial node:
//
//   y[n][th] = + mk[n] wt u                                   (mass_matrix.py:145,169)
//              - (dy/2) w_j ct sum_i' S[i][i'] u[i'][j]       (interior_convection_matrix.py:179-193)
//              - (dx/2) w_i st sum_j' S[j][j'] u[i][j']       (.. :223-237)
//              + face fluxes (outflow: own trace, inflow: neighbour trace through
//                the spatial / angular mortars)               (boundary_convection_matrix.py:284-430)
//              - ms[n] ws[n * wstride]        when ws != nullptr (scattering sum computed
//                                             by the caller into shared memory)
//
// us: the column's slab in shared memory, row stride ustride, already offset to
// the column (us[n * ustride + th]); ws likewise, already offset by th.
// N > 0: nx == ny == N at compile time; N == 0: any nx, ny <= RT_MAXN.
template <int N>
__device__ __forceinline__ void stream_column_theta(
    const DevMesh& m, const int c, const int th, const int nx, const int ny, const int sg,
    const int ntp, const int64_t tp0, const double* __restrict__ us, const int ustride,
    const double* __restrict__ ws, const int wstride, const unsigned terms,
    const int interior_only, const double* __restrict__ x, const double* __restrict__ ghost,
    double* __restrict__ y) {
    constexpr int NF = N > 0 ? N : RT_MAXN;
    const double* wx = c_w + nx * RT_MAXN;
    const double* wy = c_w + ny * RT_MAXN;
    const double* Sx = c_s + nx * RT_MAXN * RT_MAXN;
    const double* Sy = c_s + ny * RT_MAXN * RT_MAXN;
    const int nn = nx * ny;
    double* yc = y + m.col_slab_off[c] + th;
    const double hdx = 0.5 * m.col_dx[c], hdy = 0.5 * m.col_dy[c];
    const int bdry = m.col_bdry[c];
    const double* __restrict__ mk = m.coef + m.col_coef_off[c];
    const int bits = m.th_bits[tp0 + th];
    const double wt = m.th_wt[tp0 + th];
    const double ct = m.th_ct[tp0 + th];
    const double st = m.th_st[tp0 + th];

    double fa[4][NF];
#pragma unroll
    for (int F = 0; F < 4; ++F)
#pragma unroll
        for (int q = 0; q < NF; ++q) fa[F][q] = 0.0;

    if (terms & RTB200_TERM_BCONV) {
        const int cell = m.th_cell[tp0 + th];
        const int r = (bits >> 8) & 0xff;
        const int nth0 = m.cell_nth[m.sig_cell0[sg] + cell];
#pragma unroll
        for (int F = 0; F < 4; ++F) {
            if ((bits >> F) & 1) continue;       // outflow for this angle
            if ((bdry >> F) & 1) continue;       // inflow through the domain boundary
            const int n0 = (F & 1) ? nx : ny;
            const double half = (F & 1) ? hdx : hdy;
            const double* wq = (F & 1) ? wx : wy;
            const double wF = m.th_wF[F * m.n_theta_pool + tp0 + th];
            for (int slot = 0; slot < 2; ++slot) {
                const int fidx = c * 8 + F * 2 + slot;
                const int64_t base = m.face_base[fidx];
                if (base == -1) continue;
                const double* __restrict__ src = base >= 0 ? x + base : ghost + (-(base + 2));
                const int rs = m.face_rstride[fidx];
                const int eo = m.face_exy[fidx];
                const int ao = m.face_acpl[fidx];
                if (eo < 0 && ao < 0) {
                    // conforming neighbour: E = diag(w), E_theta = diag(wF)
                    const double f = half * wF;
#pragma unroll
                    for (int q = 0; q < NF; ++q)
                        if (q < n0) fa[F][q] = f * wq[q] * src[(int64_t)q * rs + th];
                    continue;
                }
                const int n1 = m.face_n1[fidx];
                for (int j1 = 0; j1 < n1; ++j1) {
                    double tmp;
                    if (ao < 0) {
                        tmp = wF * src[(int64_t)j1 * rs + th];
                    } else {
                        const int32_t* __restrict__ acp = m.acpl_pool + ao + cell * 6;
                        tmp = 0.0;
                        for (int k = 0; k < 2; ++k) {
                            const int n1t = acp[k * 3 + 2];
                            if (n1t == 0) continue;
                            const double* __restrict__ eth = m.eth_pool + acp[k * 3] + r;
                            const double* __restrict__ tr = src + (int64_t)j1 * rs + acp[k * 3 + 1];
                            for (int a = 0; a < n1t; ++a) tmp = fma(eth[a * nth0], tr[a], tmp);
                        }
                    }
                    tmp *= half;
                    if (eo < 0) {
#pragma unroll
                        for (int q = 0; q < NF; ++q)
                            if (q == j1) fa[F][q] = fma(wq[q], tmp, fa[F][q]);
                    } else {
                        const double* __restrict__ E = m.exy_pool + eo;
#pragma unroll
                        for (int q = 0; q < NF; ++q)
                            if (q < n0) fa[F][q] = fma(E[j1 * n0 + q], tmp, fa[F][q]);
                    }
                }
            }
        }
    }
    const double wF0 = m.th_wF[0 * m.n_theta_pool + tp0 + th];
    const double wF1 = m.th_wF[1 * m.n_theta_pool + tp0 + th];
    const double wF2 = m.th_wF[2 * m.n_theta_pool + tp0 + th];
    const double wF3 = m.th_wF[3 * m.n_theta_pool + tp0 + th];

#pragma unroll
    for (int i = 0; i < NF; ++i) {
        if (i >= nx) break;
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            if (j >= ny) break;
            const int n = i * ny + j;
            const double u0 = us[n * ustride + th];
            double v = 0.0;
            if (terms & RTB200_TERM_MASS) v = mk[n] * wt * u0;
            if (terms & RTB200_TERM_ICONV) {
                double sx = 0.0, sy = 0.0;
#pragma unroll
                for (int k = 0; k < NF; ++k)
                    if (k < nx) sx = fma(Sx[i * RT_MAXN + k], us[(k * ny + j) * ustride + th], sx);
#pragma unroll
                for (int k = 0; k < NF; ++k)
                    if (k < ny) sy = fma(Sy[j * RT_MAXN + k], us[(i * ny + k) * ustride + th], sy);
                v -= hdy * wy[j] * ct * sx + hdx * wx[i] * st * sy;
            }
            if (terms & RTB200_TERM_BCONV) {
                if (i == nx - 1) v += (bits & 1) ? hdy * wy[j] * wF0 * u0 : fa[0][j];
                if (j == ny - 1) v += (bits & 2) ? hdx * wx[i] * wF1 * u0 : fa[1][i];
                if (i == 0) v += (bits & 4) ? hdy * wy[j] * wF2 * u0 : fa[2][j];
                if (j == 0) v += (bits & 8) ? hdx * wx[i] * wF3 * u0 : fa[3][i];
            }
            if (ws) v = fma(-mk[nn + n], ws[n * wstride], v);
            const bool masked = interior_only && dof_masked(i, j, nx, ny, bdry, bits);
            yc[(int64_t)n * ntp] = masked ? 0.0 : v;
        }
    }
}


// Same terms for ONE node (i, j) of column c at angular index th (< nt): the item of the
// node-parallel streaming phase (k_apply<N, TH, true>).  A (column, angle) pair is spread
// over nx * ny threads instead of one, so a group of a few small hp columns fills the
// block; the price is that the threads of a mortar face each recompute the neighbour's
// angular-mortared trace values they need (the tables are L1-resident).  Operation order
// per output entry is the one of stream_column_theta: results are bit-identical.
template <int N>
__device__ __forceinline__ void stream_node_theta(
    const DevMesh& m, const int c, const int th, const int i, const int j, const int nx, const int ny,
    const int sg, const int ntp, const int64_t tp0, const double* __restrict__ us, const int ustride,
    const unsigned terms, const int interior_only, const double* __restrict__ x,
    const double* __restrict__ ghost, double* __restrict__ y) {
    constexpr int NF = N > 0 ? N : RT_MAXN;
    const double* wx = c_w + nx * RT_MAXN;
    const double* wy = c_w + ny * RT_MAXN;
    const double* Sx = c_s + nx * RT_MAXN * RT_MAXN;
    const double* Sy = c_s + ny * RT_MAXN * RT_MAXN;
    const int n = i * ny + j;
    const double hdx = 0.5 * m.col_dx[c], hdy = 0.5 * m.col_dy[c];
    const int bdry = m.col_bdry[c];
    const double* __restrict__ mk = m.coef + m.col_coef_off[c];
    const int bits = m.th_bits[tp0 + th];
    const double u0 = us[n * ustride + th];
    double v = 0.0;
    if (terms & RTB200_TERM_MASS) v = mk[n] * m.th_wt[tp0 + th] * u0;
    if (terms & RTB200_TERM_ICONV) {
        const double ct = m.th_ct[tp0 + th], st = m.th_st[tp0 + th];
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int k = 0; k < NF; ++k)
            if (k < nx) sx = fma(Sx[i * RT_MAXN + k], us[(k * ny + j) * ustride + th], sx);
#pragma unroll
        for (int k = 0; k < NF; ++k)
            if (k < ny) sy = fma(Sy[j * RT_MAXN + k], us[(i * ny + k) * ustride + th], sy);
        v -= hdy * wy[j] * ct * sx + hdx * wx[i] * st * sy;
    }
    if (terms & RTB200_TERM_BCONV) {
        const bool on_face[4] = {i == nx - 1, j == ny - 1, i == 0, j == 0};
#pragma unroll
        for (int F = 0; F < 4; ++F) {
            if (!on_face[F]) continue;
            const int q = (F & 1) ? i : j;               // position along the face
            const double half = (F & 1) ? hdx : hdy;
            const double* wq = (F & 1) ? wx : wy;
            const double wF = m.th_wF[F * m.n_theta_pool + tp0 + th];
            if ((bits >> F) & 1) {                       // outflow: own trace
                v += half * wq[q] * wF * u0;
                continue;
            }
            if ((bdry >> F) & 1) continue;               // inflow through the domain boundary
            const int n0 = (F & 1) ? nx : ny;
            const int cell = m.th_cell[tp0 + th];
            const int r = (bits >> 8) & 0xff;
            const int nth0 = m.cell_nth[m.sig_cell0[sg] + cell];
            double fa = 0.0;
            for (int slot = 0; slot < 2; ++slot) {
                const int fidx = c * 8 + F * 2 + slot;
                const int64_t base = m.face_base[fidx];
                if (base == -1) continue;
                const double* __restrict__ src = base >= 0 ? x + base : ghost + (-(base + 2));
                const int rs = m.face_rstride[fidx];
                const int eo = m.face_exy[fidx];
                const int ao = m.face_acpl[fidx];
                if (eo < 0 && ao < 0) {                  // conforming neighbour
                    fa = (half * wF) * wq[q] * src[(int64_t)q * rs + th];
                    continue;
                }
                const int n1 = m.face_n1[fidx];
                for (int j1 = 0; j1 < n1; ++j1) {
                    if (eo < 0 && j1 != q) continue;     // E = diag(w): one term
                    double tmp;
                    if (ao < 0) {
                        tmp = wF * src[(int64_t)j1 * rs + th];
                    } else {
                        const int32_t* __restrict__ acp = m.acpl_pool + ao + cell * 6;
                        tmp = 0.0;
                        for (int k = 0; k < 2; ++k) {
                            const int n1t = acp[k * 3 + 2];
                            if (n1t == 0) continue;
                            const double* __restrict__ eth = m.eth_pool + acp[k * 3] + r;
                            const double* __restrict__ tr = src + (int64_t)j1 * rs + acp[k * 3 + 1];
                            for (int a = 0; a < n1t; ++a) tmp = fma(eth[a * nth0], tr[a], tmp);
                        }
                    }
                    tmp *= half;
                    fa = fma(eo < 0 ? wq[q] : m.exy_pool[eo + j1 * n0 + q], tmp, fa);
                }
            }
            v += fa;
        }
    }
    const bool masked = interior_only && dof_masked(i, j, nx, ny, bdry, bits);
    y[m.col_slab_off[c] + th + (int64_t)n * ntp] = masked ? 0.0 : v;
}
