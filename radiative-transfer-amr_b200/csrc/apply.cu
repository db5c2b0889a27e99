// Matrix-free transport-operator apply on slab-order vectors (sm_100a).
//
//   y = [ +BC - IC + M - S ] x           (reference: src/rt/solve.py:52)
//
// One thread block owns one spatial column.  The column's slab
// u[node][theta] (node = i*ny + j, theta over all angular cells of the column)
// is staged once in shared memory; each thread then owns one angular DoF
// "theta" and produces y[node][theta] for every spatial node of the column:
//
//   scattering  -ms[node] * sum_theta' G[theta][theta'] u[node][theta']
//               (dense in angle per spatial node; preconditioner_matrix.py:154-248)
//   mass        +mk[node] * wt[theta] * u[node][theta]          (mass_matrix.py:145,169)
//   volume      -(dy/2) w_j ct[theta] sum_i' S[i][i'] u[i'][j]
//               -(dx/2) w_i st[theta] sum_j' S[j][j'] u[i][j']
//               (interior_convection_matrix.py:179-193,223-237)
//   faces       outflow: +(d/2) w_q wF[F][theta] u[face node]
//               inflow : +(d/2) sum_j E[j][q] sum_a Eth[a][r] u'[j][theta'(a)]
//               read from the neighbour column's trace
//               (boundary_convection_matrix.py:284-430)
//
// x is read once from DRAM (neighbour traces hit L2: columns are launched in
// memory order) and y is written once: 16 B/DoF of algorithmic traffic.
// Lanes of a warp walk consecutive theta, so slab rows are read and written
// fully coalesced and shared-memory accesses are conflict free; the u values
// of the scattering contraction are warp-wide broadcasts.
#include "rtb200_internal.cuh"

#include <algorithm>
#include <map>

__constant__ double c_w[(RT_MAXN + 1) * RT_MAXN];
__constant__ double c_s[(RT_MAXN + 1) * RT_MAXN * RT_MAXN];

// --------------------------------------------------------------------------- //
// apply kernel.  N > 0: nx == ny == N known at compile time, loops unrolled and
// accumulators in registers.  N == 0: any nx, ny (run-time loops).  RC = number
// of spatial nodes whose accumulators are carried through one pass over G.
// --------------------------------------------------------------------------- //
template <int N, int RC>
__global__ void __launch_bounds__(256)
k_apply(DevMesh m, unsigned terms, int interior_only, const double* __restrict__ x,
        const double* __restrict__ ghost, double* __restrict__ y,
        const int32_t* __restrict__ cols) {
    extern __shared__ __align__(16) double u_s[];
    constexpr int NF = N > 0 ? N : RT_MAXN;

    const int c = cols[blockIdx.x];
    const int nx = N > 0 ? N : m.col_nx[c];
    const int ny = N > 0 ? N : m.col_ny[c];
    const int nn = nx * ny;
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg];
    const int ntp = m.sig_ntp[sg];
    const int64_t soff = m.col_slab_off[c];
    const int64_t tp0 = m.sig_th0[sg];
    const double hdx = 0.5 * m.col_dx[c], hdy = 0.5 * m.col_dy[c];
    const int bdry = m.col_bdry[c];
    const double* __restrict__ mk = m.coef + m.col_coef_off[c];
    const double* __restrict__ ms = mk + nn;
    const double* __restrict__ Gt = m.g_pool + m.sig_g_off[sg];

    {   // stage the slab: 16-byte vector copies (slab offsets and sizes are even)
        const double2* __restrict__ src = reinterpret_cast<const double2*>(x + soff);
        double2* dst = reinterpret_cast<double2*>(u_s);
        const int n2 = (nn * ntp) >> 1;
        for (int e = threadIdx.x; e < n2; e += blockDim.x) dst[e] = src[e];
    }
    __syncthreads();

    const double* wx = c_w + nx * RT_MAXN;
    const double* wy = c_w + ny * RT_MAXN;
    const double* Sx = c_s + nx * RT_MAXN * RT_MAXN;
    const double* Sy = c_s + ny * RT_MAXN * RT_MAXN;

    for (int th = threadIdx.x; th < ntp; th += blockDim.x) {
        double* yc = y + soff + th;
        if (th >= nt) {                       // pad column of the slab stays zero
            for (int n = 0; n < nn; ++n) yc[(int64_t)n * ntp] = 0.0;
            continue;
        }
        const int bits = m.th_bits[tp0 + th];
        const double wt = m.th_wt[tp0 + th];
        const double ct = m.th_ct[tp0 + th];
        const double st = m.th_st[tp0 + th];

        // ---- inflow face fluxes fa[F][q], q = node index along the face ---- //
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
                for (int slot = 0; slot < 2; ++slot) {
                    const int fidx = c * 8 + F * 2 + slot;
                    const int64_t base = m.face_base[fidx];
                    if (base == -1) continue;
                    const double* __restrict__ src = base >= 0 ? x + base : ghost + (-(base + 2));
                    const int rs = m.face_rstride[fidx];
                    const int n1 = m.face_n1[fidx];
                    const double* __restrict__ E = m.exy_pool + m.face_exy[fidx];
                    const int32_t* __restrict__ acp = m.acpl_pool + m.face_acpl[fidx] + cell * 6;
                    for (int j1 = 0; j1 < n1; ++j1) {
                        double tmp = 0.0;
                        for (int k = 0; k < 2; ++k) {
                            const int n1t = acp[k * 3 + 2];
                            if (n1t == 0) continue;
                            const double* __restrict__ eth = m.eth_pool + acp[k * 3] + r;
                            const double* __restrict__ tr = src + (int64_t)j1 * rs + acp[k * 3 + 1];
                            for (int a = 0; a < n1t; ++a) tmp = fma(eth[a * nth0], tr[a], tmp);
                        }
                        tmp *= half;
#pragma unroll
                        for (int q = 0; q < NF; ++q)
                            if (q < n0) fa[F][q] = fma(E[j1 * n0 + q], tmp, fa[F][q]);
                    }
                }
            }
        }
        const double wF0 = m.th_wF[0 * m.n_theta_pool + tp0 + th];
        const double wF1 = m.th_wF[1 * m.n_theta_pool + tp0 + th];
        const double wF2 = m.th_wF[2 * m.n_theta_pool + tp0 + th];
        const double wF3 = m.th_wF[3 * m.n_theta_pool + tp0 + th];

        // ---- nodes, RC at a time ------------------------------------------ //
        for (int nb = 0; nb < nn; nb += RC) {
            double acc[RC];
#pragma unroll
            for (int r = 0; r < RC; ++r) acc[r] = 0.0;

            if (terms & RTB200_TERM_SCAT) {
                // acc[r] = sum_tp G^T[tp][th] * u[nb + r][tp]; two tp per step so the
                // u operand is one 16-byte broadcast load.
                const double* __restrict__ g = Gt + th;
                for (int tp = 0; tp < ntp; tp += 2) {
                    const double g0 = g[(int64_t)tp * ntp];
                    const double g1 = g[(int64_t)(tp + 1) * ntp];
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        const int n = (nb + r < nn) ? nb + r : nn - 1;
                        const double2 uu = *reinterpret_cast<const double2*>(&u_s[n * ntp + tp]);
                        acc[r] = fma(g0, uu.x, acc[r]);
                        acc[r] = fma(g1, uu.y, acc[r]);
                    }
                }
            }

#pragma unroll
            for (int r = 0; r < RC; ++r) {
                const int n = nb + r;
                if (n >= nn) break;
                const int i = n / ny, j = n - i * ny;
                const double u0 = u_s[n * ntp + th];
                double v = (terms & RTB200_TERM_SCAT) ? -ms[n] * acc[r] : 0.0;
                if (terms & RTB200_TERM_MASS) v = fma(mk[n] * wt, u0, v);
                if (terms & RTB200_TERM_ICONV) {
                    double sx = 0.0, sy = 0.0;
#pragma unroll
                    for (int k = 0; k < NF; ++k)
                        if (k < nx) sx = fma(Sx[i * RT_MAXN + k], u_s[(k * ny + j) * ntp + th], sx);
#pragma unroll
                    for (int k = 0; k < NF; ++k)
                        if (k < ny) sy = fma(Sy[j * RT_MAXN + k], u_s[(i * ny + k) * ntp + th], sy);
                    v -= hdy * wy[j] * ct * sx + hdx * wx[i] * st * sy;
                }
                bool masked = false;
                if (i == nx - 1) {
                    if (terms & RTB200_TERM_BCONV)
                        v += (bits & 1) ? hdy * wy[j] * wF0 * u0 : fa[0][j];
                    masked |= (bdry & 1) && (bits & 16);
                }
                if (j == ny - 1) {
                    if (terms & RTB200_TERM_BCONV)
                        v += (bits & 2) ? hdx * wx[i] * wF1 * u0 : fa[1][i];
                    masked |= (bdry & 2) && (bits & 32);
                }
                if (i == 0) {
                    if (terms & RTB200_TERM_BCONV)
                        v += (bits & 4) ? hdy * wy[j] * wF2 * u0 : fa[2][j];
                    masked |= (bdry & 4) && (bits & 64);
                }
                if (j == 0) {
                    if (terms & RTB200_TERM_BCONV)
                        v += (bits & 8) ? hdx * wx[i] * wF3 * u0 : fa[3][i];
                    masked |= (bdry & 8) && (bits & 128);
                }
                yc[(int64_t)n * ntp] = (interior_only && masked) ? 0.0 : v;
            }
        }
    }
}

// --------------------------------------------------------------------------- //
// layout conversion and masking (one block per column)
// --------------------------------------------------------------------------- //
__global__ void k_to_slab(DevMesh m, const double* __restrict__ xd, double* __restrict__ xs) {
    const int c = blockIdx.x;
    const int nn = m.col_nx[c] * m.col_ny[c];
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg], ntp = m.sig_ntp[sg];
    const int64_t soff = m.col_slab_off[c];
    int64_t doff = m.col_dof_off[c];
    const int cell0 = m.sig_cell0[sg], ncell = m.sig_ncell[sg];
    for (int ci = 0; ci < ncell; ++ci) {
        const int nth = m.cell_nth[cell0 + ci], tho = m.cell_thoff[cell0 + ci];
        const int cnt = nn * nth;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int node = e / nth, a = e - node * nth;
            xs[soff + (int64_t)node * ntp + tho + a] = xd[doff + e];
        }
        doff += cnt;
    }
    if (ntp > nt)
        for (int node = threadIdx.x; node < nn; node += blockDim.x)
            xs[soff + (int64_t)node * ntp + nt] = 0.0;
}

__global__ void k_from_slab(DevMesh m, const double* __restrict__ xs, double* __restrict__ xd) {
    const int c = blockIdx.x;
    const int nn = m.col_nx[c] * m.col_ny[c];
    const int sg = m.col_sig[c];
    const int ntp = m.sig_ntp[sg];
    const int64_t soff = m.col_slab_off[c];
    int64_t doff = m.col_dof_off[c];
    const int cell0 = m.sig_cell0[sg], ncell = m.sig_ncell[sg];
    for (int ci = 0; ci < ncell; ++ci) {
        const int nth = m.cell_nth[cell0 + ci], tho = m.cell_thoff[cell0 + ci];
        const int cnt = nn * nth;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int node = e / nth, a = e - node * nth;
            xd[doff + e] = xs[soff + (int64_t)node * ntp + tho + a];
        }
        doff += cnt;
    }
}

// keep == 1: zero inflow-boundary entries; keep == 2: zero everything else.
__global__ void k_mask(DevMesh m, double* __restrict__ xs, int keep) {
    const int c = blockIdx.x;
    const int nx = m.col_nx[c], ny = m.col_ny[c];
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg], ntp = m.sig_ntp[sg];
    const int64_t soff = m.col_slab_off[c];
    const int64_t tp0 = m.sig_th0[sg];
    const int bdry = m.col_bdry[c];
    const int total = nx * ny * ntp;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int node = e / ntp, th = e - node * ntp;
        if (th >= nt) { xs[soff + e] = 0.0; continue; }
        const int i = node / ny, j = node - i * ny;
        const int bits = m.th_bits[tp0 + th];
        const bool masked = ((i == nx - 1) && (bdry & 1) && (bits & 16)) ||
                            ((j == ny - 1) && (bdry & 2) && (bits & 32)) ||
                            ((i == 0) && (bdry & 4) && (bits & 64)) ||
                            ((j == 0) && (bdry & 8) && (bits & 128));
        if ((keep == 1 && masked) || (keep == 2 && !masked)) xs[soff + e] = 0.0;
    }
}

// --------------------------------------------------------------------------- //
// host side
// --------------------------------------------------------------------------- //
template <int N, int RC>
static int launch_bucket(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only,
                         const double* x, double* y, cudaStream_t st) {
    auto kern = k_apply<N, RC>;
    if (b.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)b.smem);
        if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply)");
    }
    int threads = ((b.max_ntp + 31) / 32) * 32;
    threads = std::max(64, std::min(256, threads));
    kern<<<b.count, threads, b.smem, st>>>(ctx->dm, terms, interior_only, x, ctx->ghost_dev, y,
                                           b.cols_dev);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rt_cuda(ctx, e, "k_apply launch");
    return RTB200_OK;
}

int rt_launch_apply(rtb200_ctx* ctx, unsigned terms, int interior_only, const double* x,
                    double* y, cudaStream_t st) {
    for (const Bucket& b : ctx->buckets) {
        int rc;
        switch (b.n) {
            case 2: rc = launch_bucket<2, 4>(ctx, b, terms, interior_only, x, y, st); break;
            case 3: rc = launch_bucket<3, 9>(ctx, b, terms, interior_only, x, y, st); break;
            case 4: rc = launch_bucket<4, 16>(ctx, b, terms, interior_only, x, y, st); break;
            case 5: rc = launch_bucket<5, 25>(ctx, b, terms, interior_only, x, y, st); break;
            case 6: rc = launch_bucket<6, 18>(ctx, b, terms, interior_only, x, y, st); break;
            case 7: rc = launch_bucket<7, 25>(ctx, b, terms, interior_only, x, y, st); break;
            case 8: rc = launch_bucket<8, 32>(ctx, b, terms, interior_only, x, y, st); break;
            default: rc = launch_bucket<0, 8>(ctx, b, terms, interior_only, x, y, st); break;
        }
        if (rc) return rc;
    }
    return RTB200_OK;
}

int rt_apply_setup(rtb200_ctx* ctx, const rtb200_flatmesh* m) {
    if (!m->lgl_w || !m->lgl_s) return rt_fail(ctx, RTB200_EINVAL, "missing LGL tables");
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_w, m->lgl_w, sizeof(double) * (RT_MAXN + 1) * RT_MAXN));
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_s, m->lgl_s,
                                    sizeof(double) * (RT_MAXN + 1) * RT_MAXN * RT_MAXN));
    // buckets by compile-time size; column order inside a bucket = memory order
    std::map<int, std::vector<int32_t>> by_n;
    std::map<int, int> max_ntp;
    std::map<int, size_t> smem;
    for (int c = 0; c < m->n_col; ++c) {
        const int nx = m->col_nx[c], ny = m->col_ny[c];
        const int n = (nx == ny && nx >= 2 && nx <= 8) ? nx : 0;
        by_n[n].push_back(c);
        const int ntp = m->sig_ntp[m->col_sig[c]];
        max_ntp[n] = std::max(max_ntp[n], ntp);
        smem[n] = std::max(smem[n], sizeof(double) * (size_t)nx * ny * ntp);
    }
    for (auto& kv : by_n) {
        Bucket b;
        b.n = kv.first;
        b.count = (int)kv.second.size();
        b.max_ntp = max_ntp[kv.first];
        b.smem = smem[kv.first];
        if (b.smem > 227 * 1024)
            return rt_fail(ctx, RTB200_EUNSUPPORTED,
                           "a column slab exceeds 227 KB of shared memory");
        RT_CUDA(ctx, cudaMalloc((void**)&b.cols_dev, sizeof(int32_t) * b.count));
        ctx->allocs.push_back(b.cols_dev);
        RT_CUDA(ctx, cudaMemcpy(b.cols_dev, kv.second.data(), sizeof(int32_t) * b.count,
                                cudaMemcpyHostToDevice));
        ctx->buckets.push_back(b);
    }
    return RTB200_OK;
}

static int check_ready(rtb200_ctx* ctx, const void* a, const void* b) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set (call rtb200_set_mesh first)");
    if (!a || !b) return rt_fail(ctx, RTB200_EINVAL, "null device vector");
    return RTB200_OK;
}

extern "C" int rtb200_apply(rtb200_ctx* ctx, unsigned terms, int interior_only,
                            const double* x_slab_dev, double* y_slab_dev, void* stream) {
    int rc = check_ready(ctx, x_slab_dev, y_slab_dev);
    if (rc) return rc;
    if (x_slab_dev == y_slab_dev) return rt_fail(ctx, RTB200_EINVAL, "apply: x and y must not alias");
    if (terms == 0 || (terms & ~RTB200_TERM_ALL)) return rt_fail(ctx, RTB200_EINVAL, "apply: bad term mask");
    cudaStream_t st = (cudaStream_t)stream;
    if ((terms & RTB200_TERM_BCONV) && ctx->nranks > 1) {
        rc = rtb200_halo_exchange(ctx, x_slab_dev, stream);
        if (rc) return rc;
    }
    return rt_launch_apply(ctx, terms, interior_only, x_slab_dev, y_slab_dev, st);
}

extern "C" int rtb200_to_slab(rtb200_ctx* ctx, const double* x_dof_dev, double* x_slab_dev,
                              void* stream) {
    int rc = check_ready(ctx, x_dof_dev, x_slab_dev);
    if (rc) return rc;
    if (ctx->dm.n_col == 0) return RTB200_OK;
    k_to_slab<<<ctx->dm.n_col, 256, 0, (cudaStream_t)stream>>>(ctx->dm, x_dof_dev, x_slab_dev);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}

extern "C" int rtb200_from_slab(rtb200_ctx* ctx, const double* x_slab_dev, double* x_dof_dev,
                                void* stream) {
    int rc = check_ready(ctx, x_dof_dev, x_slab_dev);
    if (rc) return rc;
    if (ctx->dm.n_col == 0) return RTB200_OK;
    k_from_slab<<<ctx->dm.n_col, 256, 0, (cudaStream_t)stream>>>(ctx->dm, x_slab_dev, x_dof_dev);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}

extern "C" int rtb200_mask(rtb200_ctx* ctx, double* x_slab_dev, int keep, void* stream) {
    int rc = check_ready(ctx, x_slab_dev, x_slab_dev);
    if (rc) return rc;
    if (keep != 1 && keep != 2) return rt_fail(ctx, RTB200_EINVAL, "mask: keep must be 1 or 2");
    if (ctx->dm.n_col == 0) return RTB200_OK;
    k_mask<<<ctx->dm.n_col, 256, 0, (cudaStream_t)stream>>>(ctx->dm, x_slab_dev, keep);
    ctx->launches++;
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}
