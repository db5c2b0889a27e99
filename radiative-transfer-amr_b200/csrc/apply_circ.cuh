// Apply kernel for columns whose angular mesh is UNIFORM (NC equal cells of P
// nodes each, NC >= 4) and whose scattering kernel G is therefore block
// circulant in the cell index: G[(K, r), (K', a)] = C[(K - K') mod NC][r, a]
// (any Phi that depends on theta - theta' only; detected numerically on the
// host, rt/tables.py:circulant_blocks).
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
// One thread block per column, one thread per angular DoF.  Phases:
//   0  U <- column slab, cp.async 8-byte copies into rows of stride SU
//      (consecutive nodes land P bank-pairs apart: conflict-free FFT gathers);
//      the column's metadata and kappa / sigma samples go to shared memory too
//   F  one thread per (node, a): gather x[K] = U[node][K P + a], forward real FFT
//      in registers, scatter the NC/2 complex slots to W[node][k][a] (16-byte stores)
//   M  one thread per (frequency slot k, output r), looping over nodes: holds
//      T[k][r][:] in registers, reads W[node][k][0..P) (broadcast among the P
//      threads of a slot, which always share a warp), writes W[node][k][r] in place;
//      the (slot group, node) iterations are split evenly over the warps
//   I  one thread per (node, a): gather the slots, inverse FFT, write the real
//      result back into W[node][K P + a]
//   S  streaming terms (mass, volume and face convection) + "- ms * W" -> y,
//      one thread per theta, coalesced 8-byte stores (stream_theta_circ below)
#pragma once

template <int N_, int NC_, int P_>
struct CircCfg {
    static constexpr int N = N_, NC = NC_, P = P_;
    static constexpr int NN = N * N, NT = NC * P, M = NC / 2;
    static constexpr int NTP = NT + (NT & 1);
    static constexpr int NWARPS = (NT + 31) / 32;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int SLOTS_PER_WARP = 32 / P;
    static constexpr int NGROUPS = (M + SLOTS_PER_WARP - 1) / SLOTS_PER_WARP;
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
    static constexpr int U_DOUBLES = (NN * SU + 1) & ~1;         // keeps W 16-byte aligned
    static constexpr size_t SMEM = sizeof(double) * (size_t)(U_DOUBLES + NN * SW);
    static constexpr int NODES_PER_PASS = THREADS / P;          // FFT passes never split a node
    static constexpr int FFT_PASSES = (NN + NODES_PER_PASS - 1) / NODES_PER_PASS;
    static constexpr int MINB = SMEM > 112 * 1024 ? 1 : (SMEM > 74 * 1024 ? 2 : (THREADS <= 160 ? 3 : 2));
};

// Inflow contribution of one face through the general mortar path (hanging
// nodes, p- or angular mismatch):
//   g[q] = (d/2) sum_j E[j][q] sum_a Eth[a][r] u'[j][theta'(a)]
// (boundary_convection_matrix.py:349-430).  Rare on the meshes this kernel
// serves, hence out of line.
template <int N>
__device__ __noinline__ void face_general(const DevMesh& m, const ColMeta& cm, const int F,
                                          const int th, const int sg, const int64_t tp0,
                                          const int bits, const double wFF, const double half,
                                          const double* __restrict__ x,
                                          const double* __restrict__ ghost, double* g /* [N] */) {
    const double* wq = c_w + N * RT_MAXN;
    const int cell = m.th_cell[tp0 + th];
    const int r = (bits >> 8) & 0xff;
    const int nth0 = m.cell_nth[m.sig_cell0[sg] + cell];
    for (int q = 0; q < N; ++q) g[q] = 0.0;
    for (int slot = 0; slot < 2; ++slot) {
        const long long base = cm.face_base[2 * F + slot];
        if (base == -1) continue;
        const double* __restrict__ src = base >= 0 ? x + base : ghost + (-(base + 2));
        const int rs = cm.rs[2 * F + slot], eo = cm.exy[2 * F + slot], ao = cm.acpl[2 * F + slot];
        const int n1 = cm.n1[2 * F + slot];
        for (int j1 = 0; j1 < n1; ++j1) {
            double tmp;
            if (ao < 0) {
                tmp = wFF * src[(int64_t)j1 * rs + th];
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
                g[j1] = fma(wq[j1], tmp, g[j1]);
            } else {
                const double* __restrict__ E = m.exy_pool + eo;
                for (int q = 0; q < N; ++q) g[q] = fma(E[j1 * N + q], tmp, g[q]);
            }
        }
    }
}

// Streaming terms of one angle of a column with nx == ny == N and a defined
// quadrant (angular level >= 2), plus the scattering combine:
//
//   y[n] = mk[n] wt u[n]                                     (mass_matrix.py:145,169)
//        - (dy/2) w_j ct sum_k S[i][k] u[k][j]               (interior_convection_matrix.py:179-193)
//        - (dx/2) w_i st sum_k S[j][k] u[i][k]               (.. :223-237)
//        + (d/2) wF[F] w_q * (outflow ? u[face node] : neighbour trace)   per face
//                                                            (boundary_convection_matrix.py:284-430)
//        - ms[n] w[n]                                        (scattering sum, phase I)
//
// For every angle exactly one x-face and one y-face are inflow faces.  All
// global loads (per-theta tables, neighbour traces) are issued up front and
// depend only on shared-memory metadata; the contractions run in two passes
// over shared memory (columns of u, then rows) so that only the N x N
// accumulators stay live in registers.
template <int N>
__device__ __forceinline__ void stream_theta_circ(
    const DevMesh& m, const ColMeta& cm, const double* __restrict__ coef_s /* mk[NN], ms[NN] */,
    const int th, const int sg, const int ntp, const int64_t tp0, const double* __restrict__ us,
    const int ustride, const double* __restrict__ ws, const int wstride, const unsigned terms,
    const int interior_only, const double* __restrict__ x, const double* __restrict__ ghost,
    double* __restrict__ y) {
    constexpr int NN = N * N;
    const double* wq = c_w + N * RT_MAXN;
    const double* S = c_s + N * RT_MAXN * RT_MAXN;
    const int64_t PT = m.n_theta_pool;
    const bool t_mass = terms & RTB200_TERM_MASS, t_ic = terms & RTB200_TERM_ICONV;
    const bool t_bc = terms & RTB200_TERM_BCONV;

    const int bits = m.th_bits[tp0 + th];
    const int Fx = (bits & 1) ? 2 : 0, Fy = (bits & 2) ? 3 : 1;        // the two inflow faces
    const double wt = m.th_wt[tp0 + th];
    const double ct = m.th_ct[tp0 + th];
    const double st = m.th_st[tp0 + th];
    const double wF_inx = m.th_wF[Fx * PT + tp0 + th], wF_outx = m.th_wF[(Fx ^ 2) * PT + tp0 + th];
    const double wF_iny = m.th_wF[Fy * PT + tp0 + th], wF_outy = m.th_wF[(Fy ^ 2) * PT + tp0 + th];
    const int bdry = cm.bdry;
    const double hdx = cm.hdx, hdy = cm.hdy;

    // ---- neighbour traces of the two inflow faces (conforming: raw values) ------ //
    double tx[N], ty[N];
    bool gen_x, gen_y;
    {
        const bool need = t_bc && !((bdry >> Fx) & 1);
        const long long base = cm.face_base[2 * Fx];
        const bool fast = need && cm.exy[2 * Fx] < 0 && cm.acpl[2 * Fx] < 0 && cm.face_base[2 * Fx + 1] == -1;
        gen_x = need && !fast;
        const double* __restrict__ src = (base >= 0 ? x + base : ghost + (-(base + 2))) + th;
        const int rs = cm.rs[2 * Fx];
#pragma unroll
        for (int q = 0; q < N; ++q) tx[q] = fast ? src[(int64_t)q * rs] : 0.0;
    }
    {
        const bool need = t_bc && !((bdry >> Fy) & 1);
        const long long base = cm.face_base[2 * Fy];
        const bool fast = need && cm.exy[2 * Fy] < 0 && cm.acpl[2 * Fy] < 0 && cm.face_base[2 * Fy + 1] == -1;
        gen_y = need && !fast;
        const double* __restrict__ src = (base >= 0 ? x + base : ghost + (-(base + 2))) + th;
        const int rs = cm.rs[2 * Fy];
#pragma unroll
        for (int q = 0; q < N; ++q) ty[q] = fast ? src[(int64_t)q * rs] : 0.0;
    }

    // ---- pass X: v[i][j] = -(dy/2) w_j ct sum_k S[i][k] u[k][j] ------------------ //
    double v[N][N];
    const double* __restrict__ ut = us + th;
    if (t_ic) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double col[N];
#pragma unroll
            for (int k = 0; k < N; ++k) col[k] = ut[(k * N + j) * ustride];
            const double cyj = -(hdy * wq[j] * ct);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double s = S[i * RT_MAXN] * col[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[i * RT_MAXN + k], col[k], s);
                v[i][j] = cyj * s;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) v[i][j] = 0.0;
    }

    // ---- inflow faces ------------------------------------------------------------ //
    if (t_bc) {
        const double cin_x = hdy * wF_inx, cin_y = hdx * wF_iny;
        if (gen_x) {
            double g[N];
            face_general<N>(m, cm, Fx, th, sg, tp0, bits, wF_inx, hdy, x, ghost, g);
#pragma unroll
            for (int q = 0; q < N; ++q) {
                if (Fx == 0) v[N - 1][q] += g[q];
                else v[0][q] += g[q];
            }
        } else if (Fx == 0) {
#pragma unroll
            for (int q = 0; q < N; ++q) v[N - 1][q] = fma(cin_x * wq[q], tx[q], v[N - 1][q]);
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q) v[0][q] = fma(cin_x * wq[q], tx[q], v[0][q]);
        }
        if (gen_y) {
            double g[N];
            face_general<N>(m, cm, Fy, th, sg, tp0, bits, wF_iny, hdx, x, ghost, g);
#pragma unroll
            for (int q = 0; q < N; ++q) {
                if (Fy == 1) v[q][N - 1] += g[q];
                else v[q][0] += g[q];
            }
        } else if (Fy == 1) {
#pragma unroll
            for (int q = 0; q < N; ++q) v[q][N - 1] = fma(cin_y * wq[q], ty[q], v[q][N - 1]);
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q) v[q][0] = fma(cin_y * wq[q], ty[q], v[q][0]);
        }
    }

    // ---- pass Y: rows of u; y-derivative, mass, outflow faces, scattering, store -- //
    const double cout_x = t_bc ? hdy * wF_outx : 0.0;     // the outflow x-face is Fx ^ 2
    const double cout_y = t_bc ? hdx * wF_outy : 0.0;
    const bool out_lo_x = Fx == 0, out_lo_y = Fy == 1;    // outflow through face 2 (i = 0) / face 3 (j = 0)
    // inflow-boundary rows removed by the interior restriction (get_masks.py:80-110)
    const bool msk = interior_only != 0;
    const bool m0 = msk && (bdry & 1) && (bits & 16), m1 = msk && (bdry & 2) && (bits & 32);
    const bool m2 = msk && (bdry & 4) && (bits & 64), m3 = msk && (bdry & 8) && (bits & 128);
    const double mw = t_mass ? wt : 0.0;
    double* yc = y + cm.slab_off + th;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double row[N];
#pragma unroll
        for (int k = 0; k < N; ++k) row[k] = ut[(i * N + k) * ustride];
        const double cxi = -(hdx * wq[i] * st);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int n = i * N + j;
            double acc = v[i][j];
            if (t_ic) {
                double s = S[j * RT_MAXN] * row[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[j * RT_MAXN + k], row[k], s);
                acc = fma(cxi, s, acc);
            }
            acc = fma(coef_s[n] * mw, row[j], acc);
            if (i == 0 && out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
            if (i == N - 1 && !out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
            if (j == 0 && out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
            if (j == N - 1 && !out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
            if (ws) acc = fma(-coef_s[NN + n], ws[n * wstride], acc);
            bool masked = false;
            if (i == N - 1) masked = masked || m0;
            if (j == N - 1) masked = masked || m1;
            if (i == 0) masked = masked || m2;
            if (j == 0) masked = masked || m3;
            yc[(int64_t)n * ntp] = masked ? 0.0 : acc;
        }
    }
}

template <int N, int NC, int P>
__global__ void __launch_bounds__(CircCfg<N, NC, P>::THREADS, CircCfg<N, NC, P>::MINB)
k_apply_circ(DevMesh m, const int32_t* __restrict__ cols, int sg, unsigned terms,
             int interior_only, const double* __restrict__ x, const double* __restrict__ ghost,
             double* __restrict__ y, const double* __restrict__ circ_tab) {
    using Cfg = CircCfg<N, NC, P>;
    constexpr int NN = Cfg::NN, NT = Cfg::NT, M = Cfg::M, NTP = Cfg::NTP;
    constexpr int SU = Cfg::SU, SW = Cfg::SW, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double smem_circ[];
    double* U = smem_circ;
    double* W = smem_circ + Cfg::U_DOUBLES;
    __shared__ ColMeta s_meta;
    __shared__ double s_coef[2 * NN];
    const int tid = threadIdx.x;
    const int64_t tp0 = m.sig_th0[sg];

    if (tid < 9) load_col_meta(m, cols[blockIdx.x], s_meta, tid);
    __syncthreads();

    // ---- phase 0: stage the slab and the coefficient samples -------------------- //
    if (tid < NT) {
        const double* __restrict__ src = x + s_meta.slab_off + tid;
        double* dst = U + tid;
#pragma unroll
        for (int n = 0; n < NN; ++n) cp_async8(dst + n * SU, src + n * NTP);
    }
    {
        const double* __restrict__ csrc = m.coef + m.col_coef_off[s_meta.c];
        for (int e = tid; e < 2 * NN; e += THREADS) cp_async8(&s_coef[e], csrc + e);
    }
    cp_async_wait_all();
    __syncthreads();

    if (terms & RTB200_TERM_SCAT) {
        // ---- phase F: forward FFTs ------------------------------------------- //
#pragma unroll 1
        for (int pass = 0; pass < Cfg::FFT_PASSES; ++pass) {
            const int node = pass * Cfg::NODES_PER_PASS + tid / P, a = tid % P;
            if (tid < Cfg::NODES_PER_PASS * P && node < NN) {
                double xs[NC], re[M], im[M];
                const double* src = U + node * SU + a;
#pragma unroll
                for (int K = 0; K < NC; ++K) xs[K] = src[K * P];
                rtfft::Real<NC>::forward(xs, re, im);
                double2* dst = reinterpret_cast<double2*>(W + node * SW + 2 * a);
#pragma unroll
                for (int k = 0; k < M; ++k) dst[k * P] = make_double2(re[k], im[k]);
            }
        }
        __syncthreads();

        // ---- phase M: per-frequency P x P block product, in place ------------- //
        {
            const int warp = tid >> 5, lane = tid & 31;
            constexpr int total = Cfg::NGROUPS * NN;
            int it = warp * total / Cfg::NWARPS;
            const int it1 = (warp + 1) * total / Cfg::NWARPS;
            const int r = lane % P, ks = lane / P;
            while (it < it1) {
                // one slot group: T[k][r][:] into registers, then the nodes of this warp
                const int grp = it / NN;
                int node = it - grp * NN;
                const int node1 = min(NN, node + (it1 - it));
                it += node1 - node;
                const int k = grp * Cfg::SLOTS_PER_WARP + ks;
                const bool active = ks < Cfg::SLOTS_PER_WARP && k < M;
                double tr[P], ti[P];
                const double2* t2 = reinterpret_cast<const double2*>(circ_tab) + (active ? (k * P + r) * P : 0);
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    const double2 t = t2[a];
                    tr[a] = t.x;
                    ti[a] = t.y;
                }
                double2* w2 = reinterpret_cast<double2*>(W + node * SW) + (active ? k * P : 0);
                // slot 0 packs the two real bins (X[0], X[NC/2]): no cross terms.  sel = 0
                // there, 1 elsewhere; the products below then cover both cases branch-free:
                //   yr = sum tr u.x - sel ti u.y ,  yi = sum sel tr u.y + (sel ? ti u.x : ti u.y)
                const bool real_slot = k == 0;
                for (; node < node1; ++node, w2 += SW / 2) {
                    double2 u[P];
#pragma unroll
                    for (int a = 0; a < P; ++a) u[a] = w2[a];
                    __syncwarp();
                    double yr = 0.0, yi = 0.0;
                    if (real_slot) {
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
                    if (active) w2[r] = make_double2(yr, yi);
                }
            }
        }
        __syncthreads();

        // ---- phase I: inverse FFTs (gather, barrier, write real rows in place) -- //
#pragma unroll 1
        for (int pass = 0; pass < Cfg::FFT_PASSES; ++pass) {
            const int node = pass * Cfg::NODES_PER_PASS + tid / P, a = tid % P;
            const bool active = tid < Cfg::NODES_PER_PASS * P && node < NN;
            double re[M], im[M];
            if (active) {
                const double2* src = reinterpret_cast<const double2*>(W + node * SW + 2 * a);
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    const double2 v = src[k * P];
                    re[k] = v.x;
                    im[k] = v.y;
                }
            }
            __syncthreads();
            if (active) {
                double xs[NC];
                rtfft::Real<NC>::inverse(re, im, xs);
                double* dst = W + node * SW + a;
#pragma unroll
                for (int K = 0; K < NC; ++K) dst[K * P] = xs[K];
            }
            __syncthreads();
        }
    }

    // ---- phase S: streaming terms + scattering combine -> y -------------------- //
    if (tid < NT) {
        stream_theta_circ<N>(m, s_meta, s_coef, tid, sg, NTP, tp0, U, SU,
                             (terms & RTB200_TERM_SCAT) ? W + tid : nullptr, SW, terms,
                             interior_only, x, ghost, y);
    } else if (tid < NTP) {                                 // pad entry of an odd row length
        double* yc = y + s_meta.slab_off + tid;
        for (int n = 0; n < NN; ++n) yc[(int64_t)n * NTP] = 0.0;
    }
}
