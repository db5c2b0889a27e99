// Apply kernel for hp-refined columns (arbitrary angular mesh: dense scattering) with
// nx == ny == N, N = 2 .. 8, and angular cells that all lie inside one quadrant
// (angular level >= 2) -- the columns of BASELINE configs[4].  Everything else
// (nx != ny, larger orders, quadrant-straddling cells) keeps k_apply (apply.cu).
//
// One thread block (8 warps, <= 85 registers: three blocks per SM) owns a GROUP of up to 7
// columns of the same class (N, angular signature), R = CG * N^2 <= 64 slab rows.  What
// changed against k_apply:
//
//   * column / face metadata, per-angle tables and the mass / scattering samples are staged
//     in shared memory once per block: the hot loops read no global table;
//   * streaming terms: a WARP owns a chunk of 32 / NPARTS consecutive angles of one column;
//     its lanes are (angle, part) pairs, a part being a contiguous range of the column's
//     nodes (k_apply gave one thread all nodes of an angle: too few items to fill a block on
//     hp columns, too many registers).  The inflow face fluxes -- the only irregular part:
//     conforming, hanging-node, p-mismatch and angular mortars -- are computed first by the
//     lanes of parts 0 (x-face) and 1 (y-face) into a warp-private shared-memory tile and
//     picked up after a __syncwarp: no block barrier between faces and streaming;
//   * the slab copies (cp.async) are waited for only after the first face fluxes, which read
//     x and the mortar tables from global memory;
//   * scattering GEMM on FP64 DMMA after the streaming stores, one n-tile (8 angles) per warp
//     and pass, all <= 8 m-tiles, accumulators added to y with a read-modify-write (L2 hits)
//     behind the block's only late barrier.
//
// Phases:   0  stage slabs (cp.async), metadata, per-angle tables
//           1  per warp task (column, angle chunk): face fluxes -> FAW, streaming terms -> y
//           2  scattering GEMM (A from shared memory, B = G through L1 from L2)
//           3  y += -ms * ACC
//
// FP64 work: 2 nt (scattering, DMMA) + ~4 N + 10 (streaming, DFMA) flop per DoF; the DMMA
// pipe at 37 TFLOP/s is what bounds this kernel on angular meshes of ~100 angles
// (DESIGN.md 4.2).
#include "rtb200_internal.cuh"

// own copy of the LGL weight / stiffness tables (this translation unit is compiled on its
// own; rt_hp_setup uploads the same data rt_apply_setup puts into apply.cu's copy)
#define c_w c_w_hp
#define c_s c_s_hp
__constant__ double c_w[(RT_MAXN + 1) * RT_MAXN];
__constant__ double c_s[(RT_MAXN + 1) * RT_MAXN * RT_MAXN];
#include "apply_common.cuh"
#include "apply_hp.h"

constexpr int HP_WARPS = HP_THREADS / 32;

// parts a (column, angle) item of the streaming phase is split into (contiguous node ranges)
__host__ __device__ constexpr int hp_nparts(int n) { return n <= 3 ? 1 : (n <= 5 ? 2 : (n == 6 ? 3 : 4)); }

// ---- inflow flux through face F of one (column, angle) ------------------------------------- //
//   g[q] = (d/2) sum_j E[j][q] sum_a Eth[a][r] u'[j][theta'(a)]
// (boundary_convection_matrix.py:349-430); conforming neighbours: E = diag(w), Eth = diag(wF).
// The neighbour's trace rows come from L2: all loads of a face slot are issued before the
// first one is used (the reference's |d n| <= 1 rule bounds the neighbour's order by N + 1;
// anything else takes the rolled loop).
template <int N>
__device__ __forceinline__ void hp_face(const DevMesh& m, const ColMeta& cm, const int F, const int th,
                                        const int bits, const double wFF, const int sg, const int64_t tp0,
                                        const double* __restrict__ x, const double* __restrict__ ghost,
                                        double (&gq)[N]) {
    constexpr int N1 = N + 1;
    const double* wq = c_w + N * RT_MAXN;
#pragma unroll
    for (int q = 0; q < N; ++q) gq[q] = 0.0;
    if ((cm.bdry >> F) & 1) return;                 // inflow through the domain boundary: masked rows
    const double half = (F & 1) ? cm.hdx : cm.hdy;
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
        const long long base = cm.face_base[2 * F + slot];
        if (base == -1) continue;
        const double* __restrict__ src = (base >= 0 ? x + base : ghost + (-(base + 2))) + th;
        const int rs = cm.rs[2 * F + slot], eo = cm.exy[2 * F + slot], ao = cm.acpl[2 * F + slot];
        if (eo < 0 && ao < 0) {
            const double f = half * wFF;
#pragma unroll
            for (int q = 0; q < N; ++q) gq[q] = f * wq[q] * src[(int64_t)q * rs];
            continue;
        }
        const int n1 = cm.n1[2 * F + slot];
        if (ao < 0 && n1 <= N1) {                   // spatial mortar only (hanging node and / or p-jump)
            double tr[N1];
#pragma unroll
            for (int j1 = 0; j1 < N1; ++j1) tr[j1] = j1 < n1 ? src[(int64_t)j1 * rs] : 0.0;
#pragma unroll
            for (int j1 = 0; j1 < N1; ++j1) tr[j1] = (wFF * tr[j1]) * half;
            if (eo < 0) {
#pragma unroll
                for (int q = 0; q < N; ++q) gq[q] = fma(wq[q], tr[q], gq[q]);
            } else {
                const double* __restrict__ E = m.exy_pool + eo;
#pragma unroll
                for (int j1 = 0; j1 < N1; ++j1) {
                    if (j1 < n1) {
#pragma unroll
                        for (int q = 0; q < N; ++q) gq[q] = fma(E[j1 * N + q], tr[j1], gq[q]);
                    }
                }
            }
            continue;
        }
        int e0 = 0, o0 = 0, c0n = 0, e1 = 0, o1 = 0, c1n = 0, nth0 = 0, r = 0;
        if (ao >= 0) {
            const int cell = m.th_cell[tp0 + th];
            const int32_t* __restrict__ acp = m.acpl_pool + ao + cell * 6;
            e0 = acp[0]; o0 = acp[1]; c0n = acp[2];
            e1 = acp[3]; o1 = acp[4]; c1n = acp[5];
            nth0 = m.cell_nth[m.sig_cell0[sg] + cell];
            r = (bits >> 8) & 0xff;
        }
        for (int j1 = 0; j1 < n1; ++j1) {
            double tmp;
            const double* __restrict__ trow = src - th + (int64_t)j1 * rs;
            if (ao < 0) {
                tmp = wFF * trow[th];
            } else {
                tmp = 0.0;
                const double* __restrict__ eth = m.eth_pool + e0 + r;
                for (int a = 0; a < c0n; ++a) tmp = fma(eth[a * nth0], trow[o0 + a], tmp);
                eth = m.eth_pool + e1 + r;
                for (int a = 0; a < c1n; ++a) tmp = fma(eth[a * nth0], trow[o1 + a], tmp);
            }
            tmp *= half;
            if (eo < 0) {
#pragma unroll
                for (int q = 0; q < N; ++q)
                    if (q == j1) gq[q] = fma(wq[q], tmp, gq[q]);
            } else {
                const double* __restrict__ E = m.exy_pool + eo + j1 * N;
#pragma unroll
                for (int q = 0; q < N; ++q) gq[q] = fma(E[q], tmp, gq[q]);
            }
        }
    }
}

// ---- streaming terms of one (column, angle), nodes [PART NN / NPARTS, (PART + 1) NN / NPARTS) -- //
//   y[n] = mk[n] wt u[n]                                     (mass_matrix.py:145,169)
//        - (dy/2) w_j ct sum_k S[i][k] u[k][j]               (interior_convection_matrix.py:179-193)
//        - (dx/2) w_i st sum_k S[j][k] u[i][k]               (.. :223-237)
//        + outflow faces: (d/2) w_q wF u   + inflow faces: FA (phase 1)
// Two passes over the slab in shared memory: rows (y-derivative, mass, outflow) and columns
// (x-derivative); every index into v[][] is a compile-time constant.
template <int N, int NPARTS, int PART>
__device__ __forceinline__ void hp_stream(const ColMeta& cm, const double* __restrict__ mk,
                                          const double* __restrict__ TH, const int bits, const int k8,
                                          const int th, const double* __restrict__ us, const int ntps,
                                          const double* __restrict__ fa, const int fstride, const int ntp,
                                          const unsigned terms, const int interior_only,
                                          double* __restrict__ y) {
    constexpr int NN = N * N, N0 = PART * NN / NPARTS, N1 = (PART + 1) * NN / NPARTS;
#define HP_MINE(i, j) ((i) * N + (j) >= N0 && (i) * N + (j) < N1)
#define HP_ROW(i) ((i) * N + N - 1 >= N0 && (i) * N < N1)
    const double* wq = c_w + N * RT_MAXN;
    const double* S = c_s + N * RT_MAXN * RT_MAXN;
    const bool t_mass = terms & RTB200_TERM_MASS, t_ic = terms & RTB200_TERM_ICONV;
    const bool t_bc = terms & RTB200_TERM_BCONV;
    const int Fx = (bits & 1) ? 2 : 0, Fy = (bits & 2) ? 3 : 1;        // the two inflow faces
    const double wt = TH[th], ct = TH[k8 + th], st = TH[2 * k8 + th];
    const double hdx = cm.hdx, hdy = cm.hdy;
    const double* __restrict__ ut = us + th;
    double v[N][N];
    {
        // outflow faces are Fx ^ 2 and Fy ^ 2: their weights wt cos(theta - F pi / 2)
        const double wox = TH[(3 + (Fx ^ 2)) * k8 + th], woy = TH[(3 + (Fy ^ 2)) * k8 + th];
        const double cout_x = t_bc ? hdy * wox : 0.0, cout_y = t_bc ? hdx * woy : 0.0;
        const bool out_lo_x = Fx == 0, out_lo_y = Fy == 1;   // outflow through face 2 (i = 0) / face 3 (j = 0)
        const double mw = t_mass ? wt : 0.0;
        const double hst = t_ic ? -(hdx * st) : 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (!HP_ROW(i)) continue;
            double row[N];
#pragma unroll
            for (int k = 0; k < N; ++k) row[k] = ut[(i * N + k) * ntps];
            const double cxi = hst * wq[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (!HP_MINE(i, j)) continue;
                double s = S[j * RT_MAXN] * row[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[j * RT_MAXN + k], row[k], s);
                double acc = cxi * s;
                acc = fma(mk[i * N + j] * mw, row[j], acc);
                if (i == 0 && out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
                if (i == N - 1 && !out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
                if (j == 0 && out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
                if (j == N - 1 && !out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
                v[i][j] = acc;
            }
        }
    }
    if (t_ic) {
        const double hct = -(hdy * ct);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double col[N];
#pragma unroll
            for (int k = 0; k < N; ++k) col[k] = ut[(k * N + j) * ntps];
            const double cyj = hct * wq[j];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (!HP_MINE(i, j)) continue;
                double s = S[i * RT_MAXN] * col[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[i * RT_MAXN + k], col[k], s);
                v[i][j] = fma(cyj, s, v[i][j]);
            }
        }
    }
    if (t_bc) {
        const double* fx = fa;                      // inflow x-face flux per j, y-face flux per i
        const double* fy = fa + N * fstride;
        if (Fx == 0) {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (HP_MINE(N - 1, q)) v[N - 1][q] += fx[q * fstride];
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (HP_MINE(0, q)) v[0][q] += fx[q * fstride];
        }
        if (Fy == 1) {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (HP_MINE(q, N - 1)) v[q][N - 1] += fy[q * fstride];
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (HP_MINE(q, 0)) v[q][0] += fy[q * fstride];
        }
    }
    const bool msk = interior_only != 0;
    const int bdry = cm.bdry;
    const bool m0 = msk && (bdry & 1) && (bits & 16), m1 = msk && (bdry & 2) && (bits & 32);
    const bool m2 = msk && (bdry & 4) && (bits & 64), m3 = msk && (bdry & 8) && (bits & 128);
    double* yc = y + cm.slab_off + th;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (!HP_MINE(i, j)) continue;
            bool masked = false;
            if (i == N - 1) masked = masked || m0;
            if (j == N - 1) masked = masked || m1;
            if (i == 0) masked = masked || m2;
            if (j == 0) masked = masked || m3;
            yc[(int64_t)(i * N + j) * ntp] = masked ? 0.0 : v[i][j];
        }
    }
#undef HP_MINE
#undef HP_ROW
}

template <int N, int NPARTS, int PART>
__device__ __forceinline__ void hp_stream_part(const int part, const ColMeta& cm, const double* mk,
                                               const double* TH, const int bits, const int k8, const int th,
                                               const double* us, const int ntps, const double* fa,
                                               const int fstride, const int ntp, const unsigned terms,
                                               const int interior_only, double* y) {
    if constexpr (PART < NPARTS) {
        if (part == PART)
            hp_stream<N, NPARTS, PART>(cm, mk, TH, bits, k8, th, us, ntps, fa, fstride, ntp, terms, interior_only, y);
        else
            hp_stream_part<N, NPARTS, PART + 1>(part, cm, mk, TH, bits, k8, th, us, ntps, fa, fstride, ntp,
                                                terms, interior_only, y);
    }
}

// ---- scattering GEMM on FP64 DMMA + epilogue, MT m-tiles x one n-tile per warp ------------- //
//   ACC[row][theta] = sum_theta' U[row][theta'] G[theta][theta'],  y[row][theta] (+)= -ms[row] ACC
// K is walked in steps of 8: the "even" DMMA uses theta' = 8 s + {0, 2, 4, 6}, the "odd" one
// 8 s + {1, 3, 5, 7}, so each lane fetches both operands of a pair with one 16-byte load.
// The G fragments are prefetched two steps ahead (they come from L2 through L1).
// Every warp of the block calls this (sync: the streaming stores of the whole block must be
// visible before the read-modify-write); warps without an n-tile in this pass are not active.
template <int MT>
__device__ __noinline__ void hp_gemm(const bool active, const bool sync, const double* __restrict__ U,
                                     const int ntps, const double* __restrict__ G, const int k8, const int nb,
                                     const int lane, const double* __restrict__ RMS,
                                     const long long* __restrict__ RY, const int* __restrict__ RI,
                                     const int* __restrict__ THb, const int R, const int ntp,
                                     const bool first_write, const bool msk, double* __restrict__ y) {
    const int lr = lane >> 2, lk = lane & 3;
    double acc[MT][2];
#pragma unroll
    for (int a = 0; a < MT; ++a) acc[a][0] = acc[a][1] = 0.0;
    if (active) {
        const double2* gp = reinterpret_cast<const double2*>(G + (int64_t)(nb * 8 + lr) * k8 + 2 * lk);
        const double* ap = U + (size_t)lr * ntps + 2 * lk;
        // G fragments four k-steps ahead (L2 latency: the resident blocks leave almost no L1)
        const int kl = k8 - 8;
        double2 b0 = gp[0], b1 = gp[min(8, kl) >> 1], b2 = gp[min(16, kl) >> 1], b3 = gp[min(24, kl) >> 1];
        for (int s8 = 0; s8 < k8; s8 += 8) {
            const double2 b4 = gp[min(s8 + 32, kl) >> 1];
#pragma unroll
            for (int a = 0; a < MT; ++a) {
                const double2 af = *reinterpret_cast<const double2*>(ap + (size_t)a * 8 * ntps + s8);
                dmma_8x8x4(acc[a][0], acc[a][1], af.x, b0.x);
                dmma_8x8x4(acc[a][0], acc[a][1], af.y, b0.y);
            }
            b0 = b1;
            b1 = b2;
            b2 = b3;
            b3 = b4;
        }
    }
    if (sync) __syncthreads();
    if (!active) return;
    // epilogue: this thread holds rows a * 8 + lr, angles nb * 8 + 2 lk + {0, 1}; the streaming
    // terms are already in y (first_write: there are none)
    const int th = nb * 8 + 2 * lk;
    if (th >= ntp) return;
    double2 old[MT];
#pragma unroll
    for (int a = 0; a < MT; ++a) {
        const int row = a * 8 + lr;
        old[a] = make_double2(0.0, 0.0);
        if (!first_write && row < R) old[a] = *reinterpret_cast<const double2*>(y + RY[row] + th);
    }
    const int2 bt = *reinterpret_cast<const int2*>(THb + th);
#pragma unroll
    for (int a = 0; a < MT; ++a) {
        const int row = a * 8 + lr;
        if (row >= R) continue;
        const double ms = RMS[row];
        const int info = RI[row];
        double s0 = -ms * acc[a][0], s1 = -ms * acc[a][1];
        if (msk && (info >> 12)) {                       // column on the domain boundary
            const int i = info & 15, j = (info >> 4) & 15, n = (info >> 8) & 15, bdry = info >> 12;
            if (dof_masked(i, j, n, n, bdry, bt.x)) s0 = 0.0;
            if (dof_masked(i, j, n, n, bdry, bt.y)) s1 = 0.0;
        }
        *reinterpret_cast<double2*>(y + RY[row] + th) = make_double2(old[a].x + s0, old[a].y + s1);
    }
}

template <int MT>
__device__ __forceinline__ void hp_gemm_mt(const int mt_n, const bool active, const bool sync, const double* U,
                                           const int ntps, const double* G, const int k8, const int nb,
                                           const int lane, const double* RMS, const long long* RY,
                                           const int* RI, const int* THb, const int R, const int ntp,
                                           const bool first_write, const bool msk, double* y) {
    if constexpr (MT >= 1) {
        if (mt_n == MT)
            hp_gemm<MT>(active, sync, U, ntps, G, k8, nb, lane, RMS, RY, RI, THb, R, ntp, first_write, msk, y);
        else
            hp_gemm_mt<MT - 1>(mt_n, active, sync, U, ntps, G, k8, nb, lane, RMS, RY, RI, THb, R, ntp,
                               first_write, msk, y);
    }
}

template <int N>
__global__ void __launch_bounds__(HP_THREADS, 3)
k_apply_hp(DevMesh m, const HpGroup* __restrict__ groups, unsigned terms, int interior_only,
           const double* __restrict__ x, const double* __restrict__ ghost, double* __restrict__ y) {
    constexpr int NN = N * N, NPARTS = hp_nparts(N);
    extern __shared__ __align__(16) double hp_s[];
    __shared__ ColMeta7 s_cm[HP_MAXCG];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const HpGroup& grp = groups[blockIdx.x];
    const int cg_n = grp.cg_n, sg = grp.sg, nt = grp.nt, ntp = grp.ntp;
    const int32_t* __restrict__ cols = grp.cols;
    const int64_t tp0 = grp.tp0;
    const int k8 = (nt + 7) & ~7, ntps = hp_ntps(ntp);
    const int R = cg_n * NN, Rp = (R + 7) & ~7;

    double* const U = hp_s;                                        // [Rp][ntps]
    double* const FA = U + (size_t)Rp * ntps;                      // [cg][2][N][ntp]
    double* const TH = FA + (size_t)cg_n * 2 * N * ntp;            // [7][k8]
    int* const THb = reinterpret_cast<int*>(TH + 7 * k8);          // [k8]
    double* const RMK = TH + 7 * k8 + k8 / 2;                      // [Rp] mass samples
    double* const RMS = RMK + Rp;                                  // [Rp] scattering samples
    long long* const RY = reinterpret_cast<long long*>(RMS + Rp);  // [Rp] offset of the row in y
    int* const RI = reinterpret_cast<int*>(RY + Rp);               // [Rp] i | j << 4 | N << 8 | bdry << 12

    // ---- phase 0: metadata, slabs, per-angle tables ------------------------------ //
    if (tid < 9 * cg_n) {
        const int cg = tid / 9;
        load_col_meta7(m, cols[cg], s_cm[cg], tid - cg * 9);
    }
    {
        const int chunks = ntp >> 1;
        for (int row = warp; row < R; row += HP_WARPS) {
            const int cg = row / NN, n = row - cg * NN;
            const double* __restrict__ src = x + grp.slab_off[cg] + (int64_t)n * ntp;
            double* dst = U + (size_t)row * ntps;
            for (int ch = lane; ch < chunks; ch += 32) cp_async16(dst + 2 * ch, src + 2 * ch);
            for (int k = ntp + lane; k < ntps; k += 32) dst[k] = 0.0;
        }
        for (int e = R * ntps + tid; e < Rp * ntps; e += HP_THREADS) U[e] = 0.0;
        for (int e = tid; e < k8; e += HP_THREADS) {
            const bool on = e < nt;
            TH[e] = on ? m.th_wt[tp0 + e] : 0.0;
            TH[k8 + e] = on ? m.th_ct[tp0 + e] : 0.0;
            TH[2 * k8 + e] = on ? m.th_st[tp0 + e] : 0.0;
#pragma unroll
            for (int F = 0; F < 4; ++F) TH[(3 + F) * k8 + e] = on ? m.th_wF[F * m.n_theta_pool + tp0 + e] : 0.0;
            THb[e] = on ? m.th_bits[tp0 + e] : 0;
        }
    }
    __syncthreads();                       // metadata visible
    for (int row = tid; row < Rp; row += HP_THREADS) {
        if (row >= R) {
            RMK[row] = RMS[row] = 0.0;
            RY[row] = 0;
            RI[row] = 0;
            continue;
        }
        const int cg = row / NN, n = row - cg * NN, i = n / N, j = n - i * N;
        const ColMeta7& cm = s_cm[cg];
        const double* __restrict__ coef = m.coef + cm.coef_off;
        RMK[row] = coef[n];
        RMS[row] = coef[NN + n];
        RY[row] = cm.cm.slab_off + (long long)n * ntp;
        RI[row] = i | (j << 4) | (N << 8) | (cm.cm.bdry << 12);
    }

    // ---- phase 1: inflow face fluxes, one thread per (column, face, angle) -> FA ---------- //
    // (reads x and the mortar tables from global memory: the slab copies are still in flight)
    const bool streaming = terms & (RTB200_TERM_MASS | RTB200_TERM_ICONV | RTB200_TERM_BCONV);
    if (terms & RTB200_TERM_BCONV) {
        const int nt32 = (nt + 31) & ~31;
        const int items = cg_n * 2 * nt32;
        for (int it = tid; it < items; it += HP_THREADS) {
            const int pair = it / nt32, th = it - pair * nt32;       // warp-uniform (column, face)
            if (th >= nt) continue;
            const int cg = pair >> 1, s = pair & 1;
            const int bits = THb[th];
            const int F = s == 0 ? ((bits & 1) ? 2 : 0) : ((bits & 2) ? 3 : 1);   // the inflow x- / y-face
            double gq[N];
            hp_face<N>(m, s_cm[cg].cm, F, th, bits, TH[(3 + F) * k8 + th], sg, tp0, x, ghost, gq);
            double* fa = FA + ((size_t)(cg * 2 + s) * N) * ntp + th;
#pragma unroll
            for (int q = 0; q < N; ++q) fa[q * ntp] = gq[q];
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- phase 2: streaming terms, one thread per (column, part, angle) -> y ------------ //
    if (streaming) {
        const int ntp32 = (ntp + 31) & ~31;
        const int items = cg_n * NPARTS * ntp32;
        for (int it = tid; it < items; it += HP_THREADS) {
            const int grp = it / ntp32, th = it - grp * ntp32;        // warp-uniform (column, part)
            if (th >= ntp) continue;
            const int cg = grp / NPARTS, part = grp - cg * NPARTS;
            const ColMeta& cm = s_cm[cg].cm;
            if (th >= nt) {                                           // pad entry of an odd row length
                double* yc = y + cm.slab_off + th;
                for (int n = part * NN / NPARTS; n < (part + 1) * NN / NPARTS; ++n) yc[(int64_t)n * ntp] = 0.0;
                continue;
            }
            hp_stream_part<N, NPARTS, 0>(part, cm, RMK + cg * NN, TH, THb[th], k8, th,
                                         U + (size_t)cg * NN * ntps, ntps,
                                         FA + (size_t)cg * 2 * N * ntp + th, ntp, ntp, terms, interior_only, y);
        }
    }
    if (!(terms & RTB200_TERM_SCAT)) return;

    // ---- phase 2 + 3: scattering GEMM (DMMA), read-modify-write epilogue ---------------- //
    const int n_tiles = k8 >> 3;
    const int mt_n = Rp >> 3;
    const double* __restrict__ G = m.g_pool + grp.g_off;              // [k8][k8]
    for (int p = 0; p * HP_WARPS < n_tiles; ++p) {
        const int nb = warp + p * HP_WARPS;
        hp_gemm_mt<HP_MT>(mt_n, nb < n_tiles, p == 0, U, ntps, G, k8, nb, lane, RMS, RY, RI, THb, R, ntp,
                          !streaming, interior_only != 0, y);
    }
}

template <int N>
static int launch_hp(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only,
                     const double* x, double* y, cudaStream_t st, int group0, int n_groups) {
    if (n_groups == 0) return RTB200_OK;
    auto kern = k_apply_hp<N>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b.smem);
    if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_hp, smem)");
    {   // shared-memory carve-out: what the resident blocks need, the rest stays L1 (G is read through it)
        const size_t per_block = b.smem + 2048;
        const int blocks = per_block * 3 <= 227 * 1024 ? 3 : (per_block * 2 <= 227 * 1024 ? 2 : 1);
        int pct = (int)((per_block * blocks * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
        pct = pct > 100 ? 100 : pct;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        if (e != cudaSuccess) return rt_cuda(ctx, e, "cudaFuncSetAttribute(k_apply_hp, carveout)");
    }
    kern<<<n_groups, HP_THREADS, b.smem, st>>>(ctx->dm, static_cast<const HpGroup*>(b.hp_groups_dev) + group0,
                                                 terms, interior_only, x, ctx->ghost_dev, y);
    ctx->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return rt_cuda(ctx, e, "k_apply_hp launch");
    return RTB200_OK;
}

int rt_hp_setup(rtb200_ctx* ctx, const rtb200_flatmesh* m) {
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_w, m->lgl_w, sizeof(double) * (RT_MAXN + 1) * RT_MAXN));
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_s, m->lgl_s, sizeof(double) * (RT_MAXN + 1) * RT_MAXN * RT_MAXN));
    return RTB200_OK;
}

int rt_hp_launch(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only, const double* x,
                 double* y, cudaStream_t st, int group0, int n_groups) {
    if (n_groups < 0) n_groups = b.n_groups - group0;
    switch (b.n) {
        case 2: return launch_hp<2>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 3: return launch_hp<3>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 4: return launch_hp<4>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 5: return launch_hp<5>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 6: return launch_hp<6>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 7: return launch_hp<7>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        case 8: return launch_hp<8>(ctx, b, terms, interior_only, x, y, st, group0, n_groups);
        default: return rt_fail(ctx, RTB200_EUNSUPPORTED, "k_apply_hp: order not compiled");
    }
}
