// Matrix-free transport-operator apply on slab-order vectors (sm_100a).
//
//   y = [ +BC - IC + M - S ] x           (reference: src/rt/solve.py:52)
//
// Terms (per column slab u[node][theta], node = i*ny + j, theta over all angular
// cells of the column):
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
#include "rtb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>

__constant__ double c_w[(RT_MAXN + 1) * RT_MAXN];
__constant__ double c_s[(RT_MAXN + 1) * RT_MAXN * RT_MAXN];

#include "apply_common.cuh"
#include "fft_real.cuh"

// --------------------------------------------------------------------------- //
// apply kernel
//
// One thread block (256 threads, 8 warps) owns a GROUP of up to CG columns of
// the same class (nx, ny, angular signature), so that they share G.  Phases:
//
//   0  stage the group's slabs in shared memory (cp.async, 16-byte chunks);
//      rows are padded to a stride = 8 (mod 16) doubles so that the DMMA
//      A-fragment loads of phase 2 are bank-conflict free.
//   1  streaming terms, one thread per (column, theta): mass, volume
//      convection and face fluxes for every spatial node -> y (global,
//      coalesced along theta).  Inflow traces of neighbour columns come from
//      global memory (L2 hits: columns are processed in memory order).
//   2  scattering as a GEMM on the FP64 tensor cores (mma.sync m8n8k4, DMMA):
//        ACC[row][theta] = sum_theta' U[row][theta'] * G[theta][theta'],
//      row = (column in group, node); M = CG*nn rows in <= 8 m-tiles per
//      chunk, warp w owns n-tiles w, w+8, ...; A fragments from shared memory,
//      B fragments straight from G in global memory (L1/L2 resident).
//   3  epilogue in the accumulator layout: y -= ms[node] * ACC (read-modify-
//      write of the y values phase 1 just wrote: L2 hits, no DRAM traffic).
//
// N > 0: nx == ny == N at compile time (phase-1 loops unrolled); N == 0: any.
// --------------------------------------------------------------------------- //
constexpr int AP_THREADS = 256;
constexpr int AP_WARPS = AP_THREADS / 32;
constexpr int AP_NTB = 3;     // n-tiles per warp per pass
constexpr int AP_MAXCG = 8;   // columns per group

// columns per group (CG*nn <= AP_ROWS rows) and m-tiles per chunk of the scattering GEMM
// (<= 8 accumulator tiles in registers; larger groups are walked in chunks)
// Short rows (<= 64 angles: hp meshes with few, low-order angular cells) get groups of up
// to 128 rows -- fuller blocks, the group's G fragments reused by more rows; long rows
// keep 64 (one chunk: G is streamed once per group).
__host__ __device__ constexpr int ap_cg(int nn, int rows = 128) { return nn >= rows ? 1 : (rows / nn > AP_MAXCG ? AP_MAXCG : rows / nn); }
__host__ __device__ constexpr int ap_mt(int n, int rows) { return n == 0 ? 8 : ((ap_cg(n * n, rows) * n * n + 7) / 8 > 8 ? 8 : (ap_cg(n * n, rows) * n * n + 7) / 8); }


constexpr int RT_HALO_RESUME_SLOTS = 4096;   // blocks of a persistent apply split around a halo exchange
#include "apply_circ.cuh"
#include "apply_circ7.cuh"
#include "apply_hp.h"

// TH threads per block: 256, or 128 for buckets of short rows (<= 64 angles: few n-tiles
// and few (column, angle) items, so half the warps of a 256-thread block would idle)
// NODE: streaming phase with one thread per (column, node, angle) instead of one per
// (column, angle) -- RTB200_HP_STREAM=node in the environment (see stream_node_theta)
template <int N, int TH, bool NODE = false>
__global__ void __launch_bounds__(TH, 512 / TH)
k_apply(DevMesh m, GroupTab gt, unsigned terms, int interior_only, const double* __restrict__ x,
        const double* __restrict__ ghost, double* __restrict__ y, int ntps) {
    extern __shared__ __align__(16) double u_s[];
    __shared__ int s_col[AP_MAXCG];
    constexpr int AP_MT = ap_mt(N, TH == 128 ? 128 : 64);   // (groups of another size are chunked)

    const int g = blockIdx.x;
    const int cg_n = gt.grp_count[g];
    if (threadIdx.x < cg_n) s_col[threadIdx.x] = gt.grp_cols[gt.grp_first[g] + threadIdx.x];
    __syncthreads();
    const int c0 = s_col[0];
    const int nx = N > 0 ? N : m.col_nx[c0];
    const int ny = N > 0 ? N : m.col_ny[c0];
    const int nn = nx * ny;
    const int sg = m.col_sig[c0];
    const int nt = m.sig_nt[sg];
    const int ntp = m.sig_ntp[sg];
    const int64_t tp0 = m.sig_th0[sg];
    const int M = cg_n * nn;

    // ---- phase 0: stage slabs ------------------------------------------------ //
    {
        const int chunks = ntp >> 1;                       // 16-byte chunks per row
        const int total = M * chunks;
        for (int e = threadIdx.x; e < total; e += TH) {
            const int row = e / chunks, ch = e - row * chunks;
            const int cg = row / nn, n = row - cg * nn;
            cp_async16(&u_s[row * ntps + 2 * ch],
                       x + m.col_slab_off[s_col[cg]] + (int64_t)n * ntp + 2 * ch);
        }
        const int padc = ntps - ntp;                       // zero the pad columns
        for (int e = threadIdx.x; e < M * padc; e += TH) {
            const int row = e / padc, k = e - row * padc;
            u_s[row * ntps + ntp + k] = 0.0;
        }
        cp_async_wait_all();
    }
    __syncthreads();

    // ---- phase 1: streaming terms, one thread per (column, theta) ------------ //
    if (NODE && (terms & (RTB200_TERM_MASS | RTB200_TERM_ICONV | RTB200_TERM_BCONV))) {
        const int per_col = nn * ntp;
        for (int item = threadIdx.x; item < cg_n * per_col; item += TH) {
            const int cg = item / per_col, rem = item - cg * per_col;
            const int n = rem / ntp, th = rem - n * ntp;       // angle fastest: coalesced stores
            const int c = s_col[cg];
            if (th >= nt) {
                y[m.col_slab_off[c] + (int64_t)n * ntp + th] = 0.0;
                continue;
            }
            const int i = n / ny;
            stream_node_theta<N>(m, c, th, i, n - i * ny, nx, ny, sg, ntp, tp0, u_s + cg * nn * ntps, ntps,
                                 terms, interior_only, x, ghost, y);
        }
    } else if (terms & (RTB200_TERM_MASS | RTB200_TERM_ICONV | RTB200_TERM_BCONV)) {
        for (int item = threadIdx.x; item < cg_n * ntp; item += TH) {
            const int cg = item / ntp, th = item - cg * ntp;
            const int c = s_col[cg];
            if (th >= nt) {
                double* yc = y + m.col_slab_off[c] + th;
                for (int n = 0; n < nn; ++n) yc[(int64_t)n * ntp] = 0.0;
                continue;
            }
            stream_column_theta<N>(m, c, th, nx, ny, sg, ntp, tp0, u_s + cg * nn * ntps, ntps,
                                   nullptr, 0, terms, interior_only, x, ghost, y);
        }
    }
    if (!(terms & RTB200_TERM_SCAT)) return;
    __syncthreads();        // phase-1 stores to y are visible to the whole block

    // ---- phase 2 + 3: scattering GEMM on DMMA, read-modify-write epilogue ----- //
    // K is walked in steps of 8: DMMA "even" uses theta' = 8s + {0,2,4,6}, DMMA
    // "odd" uses 8s + {1,3,5,7}, so that every lane fetches its two operands of a
    // step pair with ONE 16-byte load (A from shared memory, B from G).
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int lr = lane >> 2, lk = lane & 3;
        const int k8 = (nt + 7) & ~7;
        const int n_tiles = (nt + 7) >> 3;
        const double* __restrict__ G = m.g_pool + m.sig_g_off[sg];   // [n_tiles*8][k8]
        const int m_tiles = (M + 7) >> 3;
        const bool first_write = !(terms & (RTB200_TERM_MASS | RTB200_TERM_ICONV | RTB200_TERM_BCONV));

        for (int mc = 0; mc < m_tiles; mc += AP_MT) {              // chunks of <= 64 rows
            const int mt_n = min(AP_MT, m_tiles - mc);
            for (int nb = warp; nb < n_tiles; nb += (TH / 32) * AP_NTB) {   // passes over my n-tiles
                double acc[AP_MT][AP_NTB][2];
#pragma unroll
                for (int a = 0; a < AP_MT; ++a)
#pragma unroll
                    for (int b = 0; b < AP_NTB; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

                const int nbv = min(AP_NTB, (n_tiles - nb + (TH / 32) - 1) / (TH / 32));   // my tiles this pass
                const double2* gp[AP_NTB];
#pragma unroll
                for (int b = 0; b < AP_NTB; ++b) {
                    const int ntile = min(nb + b * (TH / 32), n_tiles - 1);
                    gp[b] = reinterpret_cast<const double2*>(G + (int64_t)(ntile * 8 + lr) * k8 + 2 * lk);
                }
                const double* ap = u_s + (mc * 8 + lr) * ntps + 2 * lk;

                double2 bf[AP_NTB];
#pragma unroll
                for (int b = 0; b < AP_NTB; ++b) bf[b] = gp[b][0];
                for (int s8 = 0; s8 < k8; s8 += 8) {
                    double2 bn[AP_NTB];
                    const int sn = (s8 + 8 < k8) ? s8 + 8 : s8;      // prefetch next step pair
#pragma unroll
                    for (int b = 0; b < AP_NTB; ++b) bn[b] = gp[b][sn >> 1];
#pragma unroll
                    for (int a = 0; a < AP_MT; ++a) {
                        if (a < mt_n) {
                            const double2 af = *reinterpret_cast<const double2*>(ap + a * 8 * ntps + s8);
#pragma unroll
                            for (int b = 0; b < AP_NTB; ++b)
                                if (b < nbv) {
                                    dmma_8x8x4(acc[a][b][0], acc[a][b][1], af.x, bf[b].x);
                                    dmma_8x8x4(acc[a][b][0], acc[a][b][1], af.y, bf[b].y);
                                }
                        }
                    }
#pragma unroll
                    for (int b = 0; b < AP_NTB; ++b) bf[b] = bn[b];
                }

                // epilogue: thread holds rows mt*8 + lr, thetas ntile*8 + 2*lk + {0, 1}
#pragma unroll
                for (int a = 0; a < AP_MT; ++a) {
                    if (a >= mt_n) break;
                    const int row = (mc + a) * 8 + lr;
                    if (row >= M) continue;
                    const int cg = row / nn, n = row - cg * nn;
                    const int c = s_col[cg];
                    const double msn = m.coef[m.col_coef_off[c] + nn + n];
                    const int bdry = m.col_bdry[c];
                    const int i = n / ny, j = n - i * ny;
                    double* yr = y + m.col_slab_off[c] + (int64_t)n * ntp;
#pragma unroll
                    for (int b = 0; b < AP_NTB; ++b) {
                        const int ntile = nb + b * (TH / 32);
                        const int th = ntile * 8 + 2 * lk;
                        if (ntile >= n_tiles || th >= ntp) continue;
                        double2 v = first_write ? make_double2(0.0, 0.0)
                                                : *reinterpret_cast<const double2*>(yr + th);
                        double s0 = -msn * acc[a][b][0], s1 = -msn * acc[a][b][1];
                        if (interior_only) {
                            const int b0 = m.th_bits[tp0 + th];
                            const int b1 = (th + 1 < nt) ? m.th_bits[tp0 + th + 1] : 0;
                            if (dof_masked(i, j, nx, ny, bdry, b0)) s0 = 0.0;
                            if (dof_masked(i, j, nx, ny, bdry, b1)) s1 = 0.0;
                        }
                        if (th + 1 >= nt) s1 = 0.0;
                        v.x += s0;
                        v.y += s1;
                        *reinterpret_cast<double2*>(yr + th) = v;
                    }
                }
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
static int ntps_for(int ntp) {
    // row stride in shared memory: >= ntp rounded up to a multiple of 8 (one DMMA
    // step pair), and = 8 (mod 16) doubles so that the 16-byte A-fragment loads of
    // a quarter warp (2 rows x 64 B) cover all 32 banks: conflict free
    int v = (ntp + 7) & ~7;
    while ((v & 15) != 8) v += 8;
    return v;
}

template <int N>
static int launch_bucket(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only,
                         const double* x, double* y, cudaStream_t st) {
    const bool small = b.max_ntp <= 64;
    auto kern = ctx->hp_node_stream ? (small ? k_apply<N, 128, true> : k_apply<N, 256, true>)
                                    : (small ? k_apply<N, 128, false> : k_apply<N, 256, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply, smem)");
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply, carveout)");
    GroupTab gt{b.grp_first_dev, b.grp_count_dev, b.cols_dev};
    kern<<<b.n_groups, small ? 128 : 256, b.smem, st>>>(ctx->dm, gt, terms, interior_only, x,
                                                 ctx->ghost_dev, y, b.ntps);
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return rt_cuda(ctx, e, "k_apply launch");
    return RTB200_OK;
}

// (N, NC, P) combinations the block-circulant kernel is compiled for: the
// benchmark mesh (5, 32, 5), the initial meshes of the reference's experiments
// ((5, 8, 3) manufactured-solution, (3, 8, 4) one-weak / two-strong, input.json
// mesh_params) and small cases for the parity tests.  Other uniform angular
// meshes run the dense kernel.
#ifdef RTB200_LAB      // make LAB=1: only what the c4 benchmark launches (fast kernel-tuning builds, never shipped)
#define RT_CIRC_COMBOS(X) X(5, 32, 5)
#else
#define RT_CIRC_COMBOS(X) X(5, 32, 5) X(5, 8, 3) X(3, 8, 4) X(4, 16, 4) X(3, 8, 3) X(4, 4, 4)
#endif

static bool circ_supported(int n, int nc, int p) {
#define X(N_, NC_, P_) if (n == N_ && nc == NC_ && p == P_) return true;
    RT_CIRC_COMBOS(X)
#undef X
    return false;
}

template <int N, int NC, int P>
static int launch_circ(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only,
                       const double* x, double* y, cudaStream_t st,
                       const int32_t* cols_dev = nullptr, int count = -1, int n_free = -1,
                       unsigned halo_epoch = 0, int resume = 0) {
    if (count < 0) { cols_dev = b.cols_dev; count = b.count; }
    if (n_free < 0) n_free = count;          // no column waits for a halo exchange
    if (count == 0) return RTB200_OK;
    cudaError_t e;
    if constexpr (NC >= 8) {
        // persistent, prefetching kernel (apply_circ7.cuh): 2 blocks per SM walk the column list
        // (bulk copies need 16-byte aligned rows: slab offsets are even, so x itself must be)
        if (ctx->circ_version >= 7 && b.circ_conforming && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            using Cfg = Circ7Cfg<N, NC, P>;
            auto kern = k_apply_circ7<N, NC, P>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
            if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_circ7, smem)");
            e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared);
            if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_circ7, carveout)");
            int per_sm = 2;                       // persistent grid: every block the SMs can hold
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::THREADS, Cfg::SMEM);
            if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
            // RTB200_HALO_SM_RESERVE (default 0) launches the overlapped multi-GPU pass on fewer blocks.
            // It was tried as a way to leave the halo exchange's kernels room next to the persistent
            // grid and measured useless (0.8725 / 0.8877 / 0.9067 ms for 0 / 4 / 8 SMs at N = 2): the
            // block scheduler spreads a smaller grid over all SMs, none is left empty.
            const int sms = (n_free < count || resume) ? std::max(1, ctx->sm_count - ctx->halo_sm_reserve) : ctx->sm_count;
            const int grid = std::min(count, std::max(per_sm, 1) * sms);
            // resume: 1 = first launch of a split apply (blocks note where they stopped),
            //         2 = second launch (blocks start where the first one stopped)
            if (resume && (!ctx->halo_resume_dev || 2 * grid > RT_HALO_RESUME_SLOTS))
                return rt_fail(ctx, RTB200_ESTATE, "no resume slots for the split persistent apply");
            if (b.circ_tab_host.size() != sizeof(CircTab7<Cfg>) / sizeof(double))
                return rt_fail(ctx, RTB200_EINVAL, "circulant table size mismatch");
            // column-schedule counters of this (column list, stream)
            unsigned* sched = nullptr;
            {
                const auto key = std::make_pair((const void*)b.cols_dev, (void*)st);
                auto itc = ctx->sched_slots.find(key);
                if (itc == ctx->sched_slots.end()) {
                    e = cudaMalloc((void**)&sched, 4 * sizeof(unsigned));
                    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaMalloc(schedule counters)");
                    ctx->allocs.push_back(sched);
                    e = cudaMemset(sched, 0, 4 * sizeof(unsigned));
                    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaMemset(schedule counters)");
                    ctx->sched_slots[key] = sched;
                } else {
                    sched = itc->second;
                }
            }
            CircTab7<Cfg> T;                      // 16 M P^2 bytes, passed in the constant bank
            memcpy(&T, b.circ_tab_host.data(), sizeof(T));
            kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(ctx->dm, cols_dev, count, b.circ_sig, terms,
                                                        interior_only, b.circ_herm ? 1 : 0, x,
                                                        ctx->ghost_dev, y, T, n_free, ctx->halo_flag_dev,
                                                        halo_epoch,
                                                        resume == 2 ? ctx->halo_resume_dev : nullptr,
                                                        resume == 1 ? ctx->halo_resume_dev : nullptr,
                                                        resume == 1 ? 8 : 0, sched);
            ctx->launches++;
            e = cudaGetLastError();
            if (e != cudaSuccess) return rt_cuda(ctx, e, "k_apply_circ7 launch");
            return RTB200_OK;
        }
    }
    using Cfg = CircCfg<N, NC, P>;
    auto kern = k_apply_circ<N, NC, P>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_circ, smem)");
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_circ, carveout)");
    kern<<<count, Cfg::THREADS, Cfg::SMEM, st>>>(ctx->dm, cols_dev, b.circ_sig, terms,
                                                   interior_only, x, ctx->ghost_dev, y,
                                                   b.circ_tab_dev);
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return rt_cuda(ctx, e, "k_apply_circ launch");
    return RTB200_OK;
}

static int launch_one_bucket(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only,
                             const double* x, double* y, cudaStream_t st,
                             const int32_t* cols_dev = nullptr, int count = -1, int n_free = -1,
                             unsigned halo_epoch = 0, int resume = 0) {
    int rc = RTB200_EUNSUPPORTED;
    if (b.circ_nc > 0) {
#define X(N_, NC_, P_)                                                                  \
    if (b.n == N_ && b.circ_nc == NC_ && b.circ_p == P_)                                \
        rc = launch_circ<N_, NC_, P_>(ctx, b, terms, interior_only, x, y, st, cols_dev, count, n_free, \
                                      halo_epoch, resume);
        RT_CIRC_COMBOS(X)
#undef X
        return rc;
    }
    if (b.hp) return rt_hp_launch(ctx, b, terms, interior_only, x, y, st);   // (parts: launch_bucket_part)
    switch (b.n) {
#ifndef RTB200_LAB
        case 2: rc = launch_bucket<2>(ctx, b, terms, interior_only, x, y, st); break;
        case 3: rc = launch_bucket<3>(ctx, b, terms, interior_only, x, y, st); break;
        case 4: rc = launch_bucket<4>(ctx, b, terms, interior_only, x, y, st); break;
        case 5: rc = launch_bucket<5>(ctx, b, terms, interior_only, x, y, st); break;
        case 6: rc = launch_bucket<6>(ctx, b, terms, interior_only, x, y, st); break;
        case 7: rc = launch_bucket<7>(ctx, b, terms, interior_only, x, y, st); break;
        case 8: rc = launch_bucket<8>(ctx, b, terms, interior_only, x, y, st); break;
#endif
        default: rc = launch_bucket<0>(ctx, b, terms, interior_only, x, y, st); break;
    }
    return rc;
}

// part 0: the whole bucket; 1: what reads no ghost trace; 2: the rest (multi-GPU overlap)
static int launch_bucket_part(rtb200_ctx* ctx, const Bucket& b, int part, unsigned terms, int interior_only,
                              const double* x, double* y, cudaStream_t st) {
    if (part == 0 || b.n_free < 0)
        return part == 1 ? RTB200_OK : launch_one_bucket(ctx, b, terms, interior_only, x, y, st);
    if (b.hp) {
        return part == 1 ? rt_hp_launch(ctx, b, terms, interior_only, x, y, st, 0, b.n_free)
                         : rt_hp_launch(ctx, b, terms, interior_only, x, y, st, b.n_free, b.n_groups - b.n_free);
    }
    if (b.circ_nc > 0) {
        return part == 1 ? launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_dev, b.n_free)
                         : launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_dev + b.n_free,
                                             b.count - b.n_free);
    }
    return part == 1 ? RTB200_OK : launch_one_bucket(ctx, b, terms, interior_only, x, y, st);
}

// The buckets of an hp mesh (one per compile-time size and kernel family) read the same x
// and write disjoint slabs of y, and most of them are too small to fill the GPU: they are
// launched on a few side streams forked from / joined to the caller's stream, largest
// bucket first, so that their tails overlap instead of adding up.
// halo_ev != nullptr: the halo exchange of this apply is in flight on another stream and
// completes with that event; every bucket launches what reads no ghost trace first.
int rt_launch_apply(rtb200_ctx* ctx, unsigned terms, int interior_only, const double* x,
                    double* y, cudaStream_t st, cudaEvent_t halo_ev) {
    const size_t nb = ctx->buckets.size();
    if (nb <= 1 || ctx->side_streams.empty()) {
        for (int part = halo_ev ? 1 : 0; part <= (halo_ev ? 2 : 0); ++part) {
            if (part == 2) RT_CUDA(ctx, cudaStreamWaitEvent(st, halo_ev, 0));
            for (const Bucket& b : ctx->buckets) {
                int rc = launch_bucket_part(ctx, b, part, terms, interior_only, x, y, st);
                if (rc) return rc;
            }
        }
        return RTB200_OK;
    }
    const size_t ns = ctx->side_streams.size();
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    for (size_t s = 0; s < std::min(ns, nb - 1); ++s)
        RT_CUDA(ctx, cudaStreamWaitEvent(ctx->side_streams[s], ctx->ev_fork, 0));
    int rc = RTB200_OK;
    for (int part = halo_ev ? 1 : 0; part <= (halo_ev ? 2 : 0) && rc == RTB200_OK; ++part) {
        if (part == 2) {
            cudaStreamWaitEvent(st, halo_ev, 0);
            for (size_t s = 0; s < std::min(ns, nb - 1); ++s) cudaStreamWaitEvent(ctx->side_streams[s], halo_ev, 0);
        }
        for (size_t i = 0; i < nb && rc == RTB200_OK; ++i) {
            const Bucket& b = ctx->buckets[ctx->bucket_order[i]];
            // slot 0 is the caller's stream, slots 1.. the side streams
            const size_t slot = i % (std::min(ns, nb - 1) + 1);
            rc = launch_bucket_part(ctx, b, part, terms, interior_only, x, y,
                                    slot == 0 ? st : ctx->side_streams[slot - 1]);
        }
    }
    for (size_t s = 0; s < std::min(ns, nb - 1); ++s) {      // always join, also after an error
        cudaEventRecord(ctx->ev_join[s], ctx->side_streams[s]);
        cudaStreamWaitEvent(st, ctx->ev_join[s], 0);
    }
    return rc;
}

// Apply with the face-trace halo exchange of a multi-GPU run.  When the mesh is one
// circulant bucket (the benchmark case) the columns that read no ghost trace are processed
// while the exchange (k_pack + grouped ncclSend / ncclRecv) runs on the communication
// stream; the few columns next to another rank's slab follow once the traces have landed.
__global__ void k_set_flag(unsigned* flag, unsigned value) {
    __threadfence();
    atomicExch(flag, value);
}

int rt_apply_with_halo(rtb200_ctx* ctx, unsigned terms, int interior_only, const double* x,
                       double* y, cudaStream_t st) {
    if ((terms & RTB200_TERM_BCONV) && !ctx->peer_rank.empty() && (ctx->nranks <= 1 || !ctx->nccl_comm))
        return rt_fail(ctx, RTB200_ESTATE,
                       "this rank's mesh reads face traces of other ranks but no communicator is "
                       "initialised (rtb200_comm_init): the ghost traces would be zero");
    if (!(terms & RTB200_TERM_BCONV) || ctx->nranks <= 1 || ctx->peer_rank.empty())
        return rt_launch_apply(ctx, terms, interior_only, x, y, st);
    int rc;
    if (!ctx->overlap_ready) {
        if (!ctx->overlap_multi) {
            rc = rtb200_halo_exchange(ctx, x, (void*)st);
            if (rc) return rc;
            return rt_launch_apply(ctx, terms, interior_only, x, y, st);
        }
        // several buckets (hp mesh): exchange on the communication stream, the ghost-free part
        // of every bucket meanwhile
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
        RT_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
        rc = rtb200_halo_exchange(ctx, x, (void*)ctx->comm_stream);
        if (rc) return rc;
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
        return rt_launch_apply(ctx, terms, interior_only, x, y, st, ctx->ev_halo);
    }
    const Bucket& b = ctx->buckets[0];
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    RT_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_fork, 0));
    rc = rtb200_halo_exchange(ctx, x, (void*)ctx->comm_stream);
    if (rc) return rc;
    if (ctx->ev_packed) {
        // The apply waits for the traces to be packed, so that it is released together with the
        // exchange's send / receive kernel, which sits on a high-priority stream and is placed
        // first: a persistent grid that has filled the device already would keep that kernel out
        // until its blocks leave (i.e. until they stop at their ghost-reading columns).  The blocks of
        // the apply that do not fit next to it start ~35 us late -- harmless with the dynamic
        // column schedule of k_apply_circ7.
        RT_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_packed, 0));
    }
    if (ctx->halo_fused && b.cols_all_dev && b.circ_conforming && b.circ_nc >= 8 && ctx->circ_version >= 7) {
        // ONE persistent pass over all columns, those that read no ghost trace first.  The
        // communication stream raises a flag after the exchange; a block that reaches its first
        // ghost-reading column with the flag up carries on (the usual case: the exchange is much
        // shorter than the interior work and the pipeline of the block stays warm), one that finds
        // it down notes its position and exits (it never spins: a persistent grid that waits can
        // keep the exchange's kernels off the SMs).  A second launch of the same grid, ordered
        // after the exchange, finishes what is left: normally every block returns at once.
        const unsigned epoch = ++ctx->halo_epoch;
        k_set_flag<<<1, 1, 0, ctx->comm_stream>>>(ctx->halo_flag_dev, epoch);
        ctx->launches++;
        RT_CUDA(ctx, cudaGetLastError());
        RT_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
        const int all = b.count_int + b.count_halo;
        // RTB200_HALO_FORCE_RESUME=1 (tests): the first launch looks for the NEXT epoch, so every block
        // stops at its first ghost-reading column and the second launch does real work
        const char* fr = getenv("RTB200_HALO_FORCE_RESUME");
        rc = launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_all_dev, all, b.count_int,
                               (fr && fr[0] == '1') ? epoch + 1 : epoch, 1);
        cudaStreamWaitEvent(st, ctx->ev_halo, 0);      // always rejoin
        if (rc) return rc;
        return launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_all_dev, all, all, epoch, 2);
    }
    RT_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
    rc = launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_int_dev, b.count_int);
    cudaStreamWaitEvent(st, ctx->ev_halo, 0);          // always rejoin
    if (rc) return rc;
    return launch_one_bucket(ctx, b, terms, interior_only, x, y, st, b.cols_halo_dev, b.count_halo);
}

int rt_apply_setup(rtb200_ctx* ctx, const rtb200_flatmesh* m) {
    if (!m->lgl_w || !m->lgl_s) return rt_fail(ctx, RTB200_EINVAL, "missing LGL tables");
    {
        int rc = rt_hp_setup(ctx, m);
        if (rc) return rc;
    }
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_w, m->lgl_w, sizeof(double) * (RT_MAXN + 1) * RT_MAXN));
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_s, m->lgl_s,
                                    sizeof(double) * (RT_MAXN + 1) * RT_MAXN * RT_MAXN));
    {
        double hc[32], hs[32];
        for (int t = 0; t < 32; ++t) {
            hc[t] = cos(2.0 * M_PI * t / 64.0);
            hs[t] = sin(2.0 * M_PI * t / 64.0);
        }
        hc[16] = 0.0;                                       // exact quarter turn
        RT_CUDA(ctx, cudaMemcpyToSymbol(c_fft_cos, hc, sizeof(hc)));
        RT_CUDA(ctx, cudaMemcpyToSymbol(c_fft_sin, hs, sizeof(hs)));
    }
    // Launch buckets.  Columns of the same class (nx, ny, signature) are packed
    // into groups that one thread block processes together (they share G).
    //   dense buckets    : one per compile-time size N (0 = generic), DMMA scattering;
    //   circulant buckets: one per (N, signature) whose angular mesh is uniform and
    //                      whose (N, NC, P) is compiled: FFT scattering.
    struct Cls { int n, nx, ny, sig; bool operator<(const Cls& o) const {
        return std::tie(n, nx, ny, sig) < std::tie(o.n, o.nx, o.ny, o.sig); } };
    std::map<Cls, std::vector<int32_t>> by_cls;
    for (int c = 0; c < m->n_col; ++c) {
        const int nx = m->col_nx[c], ny = m->col_ny[c];
        const int n = (nx == ny && nx >= 2 && nx <= 8) ? nx : 0;
        by_cls[Cls{n, nx, ny, m->col_sig[c]}].push_back(c);
    }
    struct Acc {
        std::vector<int32_t> cols, first, count;
        int max_ntp = 0;
        size_t rows_max = 0;
        size_t pc_rows_nt = 0;
        int pc_nn = 0;
        size_t hp_smem = 0;
        std::vector<int> sigs;
        void note(int nn, int nt, int sig) {
            pc_rows_nt = std::max(pc_rows_nt, (size_t)nn * nt);
            pc_nn = std::max(pc_nn, nn);
            if (std::find(sigs.begin(), sigs.end(), sig) == sigs.end()) sigs.push_back(sig);
        }
    };
    // key: (n, -1, 0) dense, (n, sig, mortar?) circulant.  Circulant columns whose faces are
    // all conforming (no hanging node, p- or angular mismatch) go to the persistent kernel
    // (apply_circ7.cuh, no mortar code at all); the others to k_apply_circ.
    auto has_mortar = [&](int c) {
        for (int F = 0; F < 4; ++F) {
            const int f = c * 8 + 2 * F;
            if (m->face_base[f + 1] != -1) return true;
            if (m->face_base[f] != -1 && (m->face_exy[f] >= 0 || m->face_acpl[f] >= 0)) return true;
        }
        return false;
    };
    std::map<std::tuple<int, int, int>, Acc> acc;
    for (auto& kv : by_cls) {
        const Cls& k = kv.first;
        const int nn = k.nx * k.ny;
        const bool circ = ctx->use_circ && m->sig_circ_off && m->sig_circ_off[k.sig] >= 0 && k.n > 0 &&
                          circ_supported(k.n, m->sig_circ_nc[k.sig], m->sig_circ_p[k.sig]);
        const int cg = circ ? 1 : ap_cg(nn, m->sig_ntp[k.sig] <= 64 ? 128 : 64);
        auto& v = kv.second;
        if (circ) {
            for (int c : v) {
                Acc& a = acc[{k.n, k.sig, has_mortar(c) ? 1 : 0}];
                a.first.push_back((int32_t)a.cols.size());
                a.count.push_back(1);
                a.cols.push_back(c);
                a.max_ntp = std::max(a.max_ntp, m->sig_ntp[k.sig]);
                a.note(nn, m->sig_nt[k.sig], k.sig);
            }
            continue;
        }
        // hp kernel (apply_hp.cuh): nx == ny == 2 .. 8, every angular cell inside one quadrant;
        // groups of up to 64 rows, shared memory <= 110 KB when possible (two blocks per SM)
        bool hp = ctx->use_hp && hp_supported_n(k.n);
        if (hp) {
            const int64_t o = m->sig_th0[k.sig];
            for (int t = 0; t < m->sig_nt[k.sig]; ++t) hp = hp && (m->th_bits[o + t] & 15) != 0;
        }
        if (hp) {
            const int nt = m->sig_nt[k.sig];
            int hcg = std::min(HP_MAXCG, HP_ROWS / nn);
            while (hcg > 1 && 8 * hp_smem_doubles(k.n, hcg, nt) > 110 * 1024) --hcg;   // two blocks per SM at least
            const size_t need = 8 * hp_smem_doubles(k.n, hcg, nt);
            if (need <= 220 * 1024) {
                Acc& a = acc[{k.n, -2, need > 110 * 1024 ? 2 : (need > 73 * 1024 ? 1 : 0)}];      // 1 / 2 / 3 blocks per SM
                for (size_t i = 0; i < v.size(); i += hcg) {
                    const int cnt = (int)std::min<size_t>(hcg, v.size() - i);
                    a.first.push_back((int32_t)a.cols.size());
                    a.count.push_back(cnt);
                    for (int t = 0; t < cnt; ++t) a.cols.push_back(v[i + t]);
                }
                a.hp_smem = std::max(a.hp_smem, need);
                a.max_ntp = std::max(a.max_ntp, m->sig_ntp[k.sig]);
                a.note(nn, nt, k.sig);
                continue;
            }
        }
        Acc& a = acc[{k.n, -1, 0}];
        for (size_t i = 0; i < v.size(); i += cg) {
            const int cnt = (int)std::min<size_t>(cg, v.size() - i);
            a.first.push_back((int32_t)a.cols.size());
            a.count.push_back(cnt);
            for (int t = 0; t < cnt; ++t) a.cols.push_back(v[i + t]);
            a.rows_max = std::max(a.rows_max, (size_t)((cnt * nn + 7) / 8) * 8);
        }
        a.max_ntp = std::max(a.max_ntp, m->sig_ntp[k.sig]);
        a.note(nn, m->sig_nt[k.sig], k.sig);
    }
    for (auto& kv : acc) {
        Acc& a = kv.second;
        Bucket b;
        b.n = std::get<0>(kv.first);
        b.count = (int)a.cols.size();
        b.n_groups = (int)a.first.size();
        b.max_ntp = a.max_ntp;
        b.pc_max_rows_nt = a.pc_rows_nt;
        b.pc_max_nn = a.pc_nn;
        b.sigs = a.sigs;
        if (std::get<1>(kv.first) >= 0) {
            const int sig = std::get<1>(kv.first);
            b.circ_conforming = std::get<2>(kv.first) == 0;
            b.circ_nc = m->sig_circ_nc[sig];
            b.circ_p = m->sig_circ_p[sig];
            b.circ_sig = sig;
            b.circ_tab_dev = ctx->circ_pool_dev + m->sig_circ_off[sig];
            const double* th = m->circ_pool + m->sig_circ_off[sig];
            b.circ_tab_host.assign(th, th + (size_t)(b.circ_nc / 2) * b.circ_p * b.circ_p * 2);
            {   // Hermitian block spectrum?  (slot 0 packs two real matrices: both symmetric)
                const int P = b.circ_p, M = b.circ_nc / 2;
                double tmax = 0.0, dmax = 0.0;
                for (int k = 0; k < M; ++k)
                    for (int r = 0; r < P; ++r)
                        for (int q = 0; q < P; ++q) {
                            const double* t = th + 2 * ((size_t)(k * P + r) * P + q);
                            const double* u = th + 2 * ((size_t)(k * P + q) * P + r);
                            tmax = std::max(tmax, std::max(std::fabs(t[0]), std::fabs(t[1])));
                            dmax = std::max(dmax, std::fabs(t[0] - u[0]));
                            dmax = std::max(dmax, k == 0 ? std::fabs(t[1] - u[1]) : std::fabs(t[1] + u[1]));
                        }
                const char* env = getenv("RTB200_CIRC_HERMITIAN");
                b.circ_herm = dmax <= 1e-13 * tmax && !(env && env[0] == '0');
            }
        } else if (std::get<1>(kv.first) == -2) {
            b.hp = true;
            b.smem = a.hp_smem;
            std::vector<HpGroup> recs(a.first.size());
            std::vector<char> reads_ghost(recs.size(), 0);
            for (size_t gi = 0; gi < recs.size(); ++gi) {
                HpGroup& r = recs[gi];
                memset(&r, 0, sizeof(r));
                r.cg_n = a.count[gi];
                const int c0 = a.cols[a.first[gi]];
                r.sg = m->col_sig[c0];
                r.nt = m->sig_nt[r.sg];
                r.ntp = m->sig_ntp[r.sg];
                r.tp0 = m->sig_th0[r.sg];
                r.g_off = m->sig_g_off[r.sg];
                for (int t = 0; t < r.cg_n; ++t) {
                    r.cols[t] = a.cols[a.first[gi] + t];
                    r.slab_off[t] = m->col_slab_off[r.cols[t]];
                    for (int f = 0; f < 8; ++f) reads_ghost[gi] |= m->face_base[r.cols[t] * 8 + f] <= -2;
                }
            }
            if (m->n_peer > 0) {          // groups that read no ghost trace first
                std::vector<HpGroup> ordered;
                ordered.reserve(recs.size());
                for (size_t gi = 0; gi < recs.size(); ++gi)
                    if (!reads_ghost[gi]) ordered.push_back(recs[gi]);
                b.n_free = (int)ordered.size();
                for (size_t gi = 0; gi < recs.size(); ++gi)
                    if (reads_ghost[gi]) ordered.push_back(recs[gi]);
                recs.swap(ordered);
            }
            void* d = nullptr;
            RT_CUDA(ctx, cudaMalloc(&d, sizeof(HpGroup) * recs.size()));
            ctx->allocs.push_back(d);
            RT_CUDA(ctx, cudaMemcpy(d, recs.data(), sizeof(HpGroup) * recs.size(), cudaMemcpyHostToDevice));
            b.hp_groups_dev = d;
        } else {
            b.ntps = ntps_for(b.max_ntp);
            b.smem = sizeof(double) * a.rows_max * (size_t)b.ntps;
            if (b.smem > 227 * 1024)
                return rt_fail(ctx, RTB200_EUNSUPPORTED, "a column slab exceeds 227 KB of shared memory");
        }
        if (m->n_peer > 0 && b.circ_nc > 0) {
            std::vector<int32_t> ci, ch;
            for (int32_t c : a.cols) {
                bool ghost = false;
                for (int f = 0; f < 8; ++f) ghost = ghost || m->face_base[c * 8 + f] <= -2;
                (ghost ? ch : ci).push_back(c);
            }
            b.n_free = (int)ci.size();
            ci.insert(ci.end(), ch.begin(), ch.end());
            a.cols.swap(ci);              // (one column per group: first / count stay valid)
        }
        auto up = [&](const std::vector<int32_t>& h, int32_t** d) -> int {
            RT_CUDA(ctx, cudaMalloc((void**)d, sizeof(int32_t) * h.size()));
            ctx->allocs.push_back(*d);
            RT_CUDA(ctx, cudaMemcpy(*d, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
            return RTB200_OK;
        };
        int rc;
        if ((rc = up(a.cols, &b.cols_dev))) return rc;
        if ((rc = up(a.first, &b.grp_first_dev))) return rc;
        if ((rc = up(a.count, &b.grp_count_dev))) return rc;
        ctx->buckets.push_back(b);
    }
    // multi-GPU: one circulant bucket -> split its columns into those that read a ghost trace
    // (face_base <= -2) and the rest, for the overlapped halo exchange of rt_apply_with_halo
    ctx->overlap_ready = false;
    bool single_persistent = false;
    if (m->n_peer > 0 && ctx->buckets.size() == 1 && ctx->buckets[0].circ_nc > 0) {
        Bucket& b = ctx->buckets[0];
        const auto& cols = acc.begin()->second.cols;
        std::vector<int32_t> ci, ch;
        for (int32_t c : cols) {
            bool ghost = false;
            for (int f = 0; f < 8; ++f) ghost = ghost || m->face_base[c * 8 + f] <= -2;
            (ghost ? ch : ci).push_back(c);
        }
        auto up1 = [&](const std::vector<int32_t>& h, int32_t** d) -> int {
            *d = nullptr;
            if (h.empty()) return RTB200_OK;
            RT_CUDA(ctx, cudaMalloc((void**)d, sizeof(int32_t) * h.size()));
            ctx->allocs.push_back(*d);
            RT_CUDA(ctx, cudaMemcpy(*d, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
            return RTB200_OK;
        };
        int rc;
        if ((rc = up1(ci, &b.cols_int_dev))) return rc;
        if ((rc = up1(ch, &b.cols_halo_dev))) return rc;
        {
            std::vector<int32_t> all(ci);
            all.insert(all.end(), ch.begin(), ch.end());
            if ((rc = up1(all, &b.cols_all_dev))) return rc;
            if (!ctx->halo_flag_dev) {
                RT_CUDA(ctx, cudaMalloc((void**)&ctx->halo_flag_dev, sizeof(unsigned)));
                ctx->allocs.push_back(ctx->halo_flag_dev);
                RT_CUDA(ctx, cudaMemset(ctx->halo_flag_dev, 0, sizeof(unsigned)));
                ctx->halo_epoch = 0;
                RT_CUDA(ctx, cudaMalloc((void**)&ctx->halo_resume_dev, sizeof(int) * RT_HALO_RESUME_SLOTS));
                ctx->allocs.push_back(ctx->halo_resume_dev);
            }
        }
        b.count_int = (int)ci.size();
        b.count_halo = (int)ch.size();
        if (!ctx->ev_packed) RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_packed, cudaEventDisableTiming));
        if (!ctx->comm_stream) {
            int prio_lo = 0, prio_hi = 0;
            RT_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            RT_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
            RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
            if (!ctx->ev_fork) RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        }
        // Default for this case (one circulant bucket = one persistent grid that fills every SM):
        // exchange FIRST, then one plain launch.  Overlapping does not pay: the exchange's send /
        // receive kernel cannot become resident next to two of the apply's blocks on an SM, so it
        // runs only when they leave -- measured on 2 GPUs (profiles/r02_ae_*): half of the blocks
        // find the halo flag down at their first ghost-reading column (`halo_stops` of
        // rtb200_apply_path) and the pass costs 0.858 ms against 0.848 ms serial (0.830 without any
        // exchange); a high-priority communication stream and SMs left free do not change that.
        // RTB200_HALO_OVERLAP=1 selects the overlapped single pass (dynamic schedule, resume launch).
        const char* env = getenv("RTB200_HALO_OVERLAP");
        ctx->overlap_ready = env && env[0] == '1';
        const char* fenv = getenv("RTB200_HALO_FUSED");
        ctx->halo_fused = !(fenv && fenv[0] == '0');
        single_persistent = true;
    }
    ctx->overlap_multi = false;
    if (m->n_peer > 0 && !ctx->overlap_ready && !single_persistent) {
        if (!ctx->comm_stream) {
            int prio_lo = 0, prio_hi = 0;
            RT_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            RT_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, prio_hi));
            RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        }
        if (!ctx->ev_fork) RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        const char* env = getenv("RTB200_HALO_OVERLAP");
        ctx->overlap_multi = !(env && env[0] == '0');
    }
    // launch order: most columns x angles first
    ctx->bucket_order.resize(ctx->buckets.size());
    for (size_t i = 0; i < ctx->buckets.size(); ++i) ctx->bucket_order[i] = (int)i;
    std::sort(ctx->bucket_order.begin(), ctx->bucket_order.end(), [&](int a, int b) {
        const Bucket &A = ctx->buckets[a], &B = ctx->buckets[b];
        return (int64_t)A.count * A.max_ntp * std::max(A.n, 1) * std::max(A.n, 1) >
               (int64_t)B.count * B.max_ntp * std::max(B.n, 1) * std::max(B.n, 1);
    });
    if (ctx->buckets.size() > 1 && ctx->side_streams.empty()) {
        constexpr int NSIDE = 5;
        for (int s = 0; s < NSIDE; ++s) {
            cudaStream_t q;
            cudaEvent_t ev;
            RT_CUDA(ctx, cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
            ctx->side_streams.push_back(q);
            RT_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->ev_join.push_back(ev);
        }
        if (!ctx->ev_fork) RT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
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
    return rt_apply_with_halo(ctx, terms, interior_only, x_slab_dev, y_slab_dev, st);
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

extern "C" int rtb200_apply_path(const rtb200_ctx* ctx, char* buf, int32_t buf_len) {
    if (!ctx || !buf || buf_len <= 0) return RTB200_EINVAL;
    if (!ctx->have_mesh) return RTB200_ESTATE;
    std::string s;
    char line[256];
    for (const Bucket& b : ctx->buckets) {
        if (b.circ_nc > 0) {
            const bool v7 = b.circ_nc >= 8 && ctx->circ_version >= 7 && b.circ_conforming;
            // blocks of a multi-GPU apply that found the halo flag down at their first ghost-reading
            // column so far (they resume in the second launch): 0 when the exchange is overlapped
            unsigned stops = 0;
            for (const auto& kv : ctx->sched_slots)
                if (kv.first.first == (const void*)b.cols_dev) {
                    unsigned h[4] = {0, 0, 0, 0};
                    if (cudaMemcpy(h, kv.second, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) stops += h[2];
                }
            snprintf(line, sizeof(line), "%s<%d,%d,%d> columns=%d groups=%d halo_stops=%u\n",
                     v7 ? "k_apply_circ7" : "k_apply_circ", b.n, b.circ_nc, b.circ_p, b.count, b.count, stops);
        } else if (b.hp) {
            snprintf(line, sizeof(line), "k_apply_hp<%d> columns=%d groups=%d smem=%zu\n", b.n, b.count,
                     b.n_groups, b.smem);
        } else {
            snprintf(line, sizeof(line), "k_apply<%d,%d> columns=%d groups=%d smem=%zu\n", b.n,
                     b.max_ntp <= 64 ? 128 : 256, b.count, b.n_groups, b.smem);
        }
        s += line;
    }
    snprintf(buf, (size_t)buf_len, "%s", s.c_str());
    return (int)ctx->buckets.size();
}

// node-wise block-Jacobi preconditioner (shares the transforms of k_apply_circ7)
#include "precond_eig.cuh"
#include "precond_circ.cuh"
