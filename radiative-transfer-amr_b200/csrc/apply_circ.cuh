// Apply kernel for columns whose angular mesh is UNIFORM (NC equal cells of P
// nodes each) and whose scattering kernel G is therefore block circulant in the
// cell index: G[(K, r), (K', a)] = C[(K - K') mod NC][r, a]   (any Phi that
// depends on theta - theta' only; detected numerically on the host,
// rt/tables.py:circulant_blocks).
//
// The dense product costs 2 NT flop per DoF (NT = NC P; 320 at the benchmark
// mesh, which makes the apply FP64-bound at ~1/3 of the HBM roofline).  Here it
// is done per spatial node as
//     FFT over the cell index (length NC, per intra-cell node a)
//  -> P x P complex block product per frequency (table T, shared by all nodes)
//  -> inverse FFT
// for ~(5 log2 NC + 4 P) flop per DoF (~45 at NC = 32, P = 5), which brings the
// whole operator back under the HBM roofline.  Same result as the dense path
// to rounding (a few 1e-16 relative).
//
// One thread block owns CPB columns (CPB = 1 when NT >= 128).  Phases:
//   0  U <- column slab(s), cp.async 8-byte copies into rows of odd-ish stride SU
//      (consecutive nodes land P bank-pairs apart: conflict-free FFT gathers)
//   F  one thread per (row, a): gather x[K] = U[row][K P + a], forward real FFT in
//      registers, scatter the NC/2 complex slots to W[row][k][a] (16-byte stores)
//   M  one thread per (frequency slot k, output r), looping over rows: holds
//      T[k][r][:] in registers, reads W[row][k][0..P) (broadcast among the P
//      threads of a slot, which always share a warp), writes W[row][k][r] in place
//   I  one thread per (row, a): gather the slots, inverse FFT, write the real
//      result back into W[row][K P + a]
//   1  streaming terms (mass, volume and face convection) + "- ms * W" -> y,
//      one thread per (column, theta), coalesced 8-byte stores
#pragma once

template <int N_, int NC_, int P_>
struct CircCfg {
    static constexpr int N = N_, NC = NC_, P = P_;
    static constexpr int NN = N * N, NT = NC * P, M = NC / 2;
    static constexpr int NTP = NT + (NT & 1);
    static constexpr int CPB = NT >= 128 ? 1 : (192 / NT > 8 ? 8 : 192 / NT);
    static constexpr int ROWS = CPB * NN;
    // middle phase: the P threads of one frequency slot share a warp
    static constexpr int SLOTS_PER_WARP = 32 / P;
    static constexpr int WARPS_PER_PART = (M + SLOTS_PER_WARP - 1) / SLOTS_PER_WARP;
    static constexpr int NW_STREAM = (CPB * NT + 31) / 32;
    static constexpr int NWARPS = (NW_STREAM + WARPS_PER_PART - 1) / WARPS_PER_PART * WARPS_PER_PART;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int PARTS = NWARPS / WARPS_PER_PART;
    // shared-memory row strides (doubles)
    static constexpr int su() {
        int v = NT;
        while ((v & 15) != (P & 15)) ++v;
        return v;
    }
    static constexpr int sw() {
        int v = NT + (NT & 1);
        while (((v / 2) & 7) != (P & 7)) v += 2;
        return v;
    }
    static constexpr int SU = su(), SW = sw();
    static constexpr int U_DOUBLES = (ROWS * SU + 1) & ~1;       // keeps W 16-byte aligned
    static constexpr size_t SMEM = sizeof(double) * (size_t)(U_DOUBLES + ROWS * SW);
    static constexpr int ROWS_PER_PASS = THREADS / P;           // FFT passes never split a row
    static constexpr int MINB = SMEM > 100 * 1024 ? 1 : (SMEM > 72 * 1024 ? 2 : 3);
};

template <int N, int NC, int P>
__global__ void __launch_bounds__(CircCfg<N, NC, P>::THREADS, CircCfg<N, NC, P>::MINB)
k_apply_circ(DevMesh m, GroupTab gt, unsigned terms, int interior_only,
             const double* __restrict__ x, const double* __restrict__ ghost,
             double* __restrict__ y, const double* __restrict__ circ_tab) {
    using Cfg = CircCfg<N, NC, P>;
    constexpr int NN = Cfg::NN, NT = Cfg::NT, M = Cfg::M, NTP = Cfg::NTP;
    constexpr int SU = Cfg::SU, SW = Cfg::SW, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double smem_circ[];
    double* U = smem_circ;
    double* W = smem_circ + Cfg::U_DOUBLES;
    __shared__ int s_col[Cfg::CPB];

    const int g = blockIdx.x;
    const int cg_n = gt.grp_count[g];
    if (threadIdx.x < cg_n) s_col[threadIdx.x] = gt.grp_cols[gt.grp_first[g] + threadIdx.x];
    __syncthreads();
    const int rows = cg_n * NN;
    const int sg = m.col_sig[s_col[0]];
    const int64_t tp0 = m.sig_th0[sg];

    // ---- phase 0: stage the slabs -------------------------------------------- //
    for (int e = threadIdx.x; e < rows * NT; e += THREADS) {
        const int row = e / NT, th = e - row * NT;
        const int cg = row / NN, n = row - cg * NN;
        cp_async8(&U[row * SU + th], x + m.col_slab_off[s_col[cg]] + (int64_t)n * NTP + th);
    }
    cp_async_wait_all();
    __syncthreads();

    if (terms & RTB200_TERM_SCAT) {
        // ---- phase F: forward FFTs ------------------------------------------- //
        for (int it = threadIdx.x; it < rows * P; it += THREADS) {
            const int row = it / P, a = it - row * P;
            double xs[NC], re[M], im[M];
            const double* src = U + row * SU + a;
#pragma unroll
            for (int K = 0; K < NC; ++K) xs[K] = src[K * P];
            rtfft::Real<NC>::forward(xs, re, im);
            double2* dst = reinterpret_cast<double2*>(W + row * SW + 2 * a);
#pragma unroll
            for (int k = 0; k < M; ++k) dst[k * P] = make_double2(re[k], im[k]);
        }
        __syncthreads();

        // ---- phase M: per-frequency P x P block product, in place ------------- //
        {
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            const int part = warp / Cfg::WARPS_PER_PART;
            const int k = (warp - part * Cfg::WARPS_PER_PART) * Cfg::SLOTS_PER_WARP + lane / P;
            const int r = lane % P;
            const bool active = lane < Cfg::SLOTS_PER_WARP * P && k < M;
            double tr[P], ti[P];
            if (active) {
                const double2* t2 = reinterpret_cast<const double2*>(circ_tab) + (k * P + r) * P;
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    const double2 t = t2[a];
                    tr[a] = t.x;
                    ti[a] = t.y;
                }
            }
            for (int row = part; row < rows; row += Cfg::PARTS) {
                double2* w2 = reinterpret_cast<double2*>(W + row * SW) + (active ? k * P : 0);
                double2 u[P];
#pragma unroll
                for (int a = 0; a < P; ++a) u[a] = w2[a];
                __syncwarp();
                if (active) {
                    double yr = 0.0, yi = 0.0;
                    if (k == 0) {                     // two real bins packed as (X[0], X[NC/2])
#pragma unroll
                        for (int a = 0; a < P; ++a) {
                            yr = fma(tr[a], u[a].x, yr);
                            yi = fma(ti[a], u[a].y, yi);
                        }
                    } else {
#pragma unroll
                        for (int a = 0; a < P; ++a) {
                            yr = fma(tr[a], u[a].x, yr);
                            yr = fma(-ti[a], u[a].y, yr);
                            yi = fma(tr[a], u[a].y, yi);
                            yi = fma(ti[a], u[a].x, yi);
                        }
                    }
                    w2[r] = make_double2(yr, yi);
                }
            }
        }
        __syncthreads();

        // ---- phase I: inverse FFTs (gather, barrier, write real rows in place) -- //
        for (int row0 = 0; row0 < rows; row0 += Cfg::ROWS_PER_PASS) {
            const int row = row0 + threadIdx.x / P, a = threadIdx.x % P;
            const bool active = threadIdx.x < Cfg::ROWS_PER_PASS * P && row < rows;
            double re[M], im[M], xs[NC];
            if (active) {
                const double2* src = reinterpret_cast<const double2*>(W + row * SW + 2 * a);
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double2 v = src[k * P];
                    re[k] = v.x;
                    im[k] = v.y;
                }
            }
            __syncthreads();
            if (active) {
                rtfft::Real<NC>::inverse(re, im, xs);
                double* dst = W + row * SW + a;
#pragma unroll
                for (int K = 0; K < NC; ++K) dst[K * P] = xs[K];
            }
            __syncthreads();
        }
    }

    // ---- phase 1: streaming terms + scattering combine -> y -------------------- //
    const bool scat = (terms & RTB200_TERM_SCAT) != 0;
    for (int item = threadIdx.x; item < cg_n * NTP; item += THREADS) {
        const int cg = item / NTP, th = item - cg * NTP;
        const int c = s_col[cg];
        if (th >= NT) {                                   // pad entry of an odd row length
            double* yc = y + m.col_slab_off[c] + th;
            for (int n = 0; n < NN; ++n) yc[(int64_t)n * NTP] = 0.0;
            continue;
        }
        stream_column_theta<N>(m, c, th, N, N, sg, NTP, tp0, U + cg * NN * SU, SU,
                               scat ? W + cg * NN * SW + th : nullptr, SW, terms, interior_only,
                               x, ghost, y);
    }
}
