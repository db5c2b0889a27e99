// Persistent, prefetching apply kernel for columns with a UNIFORM angular mesh
// (NC >= 8 equal cells of P nodes; block-circulant scattering, see apply_circ.cuh
// for the algorithm and the table conventions) whose faces are all conforming
// (columns with a mortar face are routed to k_apply_circ by the host).
//
// What changed against k_apply_circ (one block per column, phases in lock step):
//   * persistent blocks (as many as the SMs hold: 2 per SM at the benchmark size) walk
//     the column list with a grid stride and PREFETCH the next column's slab +
//     coefficient samples with bulk asynchronous copies (cp.async.bulk, one per slab
//     row, completing on an mbarrier) into a second buffer while the current column is
//     processed: the HBM stream and the FP64 phases overlap by construction;
//   * twice the threads per column at <= 96 registers: every length-NC real FFT is
//     split over the two lanes (l, l ^ 16) of a warp -- each transforms the even or
//     the odd samples (length NC/2, in registers), the last radix-2 stage swaps NC/4
//     doubles with shuffles -- and the streaming terms of an angle are split over
//     two threads (first / second half of the spatial nodes);
//   * a warp owns whole spatial nodes in the FFT phases (the inverse transform writes
//     its rows in place with only a __syncwarp); the block product runs with one lane
//     per node and a warp-uniform frequency slot, two slots interleaved;
//   * the last warp is the producer (bulk copies, column metadata two columns ahead in
//     a 3-slot ring) and stays out of the FFT / block-product phases;
//   * per-angle tables (weights, wt cos, wt sin, upwind bits) and the block spectrum
//     are staged in shared memory once per block.
//
// Shared memory per block at (N, NC, P) = (5, 32, 5): 2 x 33.2 KB (U) + 34.0 KB (W)
// + 6.4 KB (block spectrum) + 4.5 KB (angle tables) + 0.8 KB = 112.1 KB -> 2 blocks / SM.
// Measurements and what bounds the kernel: DESIGN.md 4.1, profiles/r01_v13_summary.md.
#pragma once
#include <utility>

template <int N_, int NC_, int P_>
struct Circ7Cfg {
    static constexpr int N = N_, NC = NC_, P = P_;
    static constexpr int NN = N * N, NT = NC * P, M = NC / 2, H = NC / 2, Q = NC / 4;
    static constexpr int NTP = NT + (NT & 1);
    static constexpr int NPW = 16 / P;                         // nodes a warp can transform (2 lanes / sequence)
    static constexpr int SC_WARPS = (NN + NPW - 1) / NPW;
    static constexpr int NTW = (NT + 31) / 32 * 32;            // S role: threads per half
    static constexpr int THREADS = SC_WARPS * 32 > 2 * NTW ? SC_WARPS * 32 : 2 * NTW;
    static constexpr int NWARPS = THREADS / 32;
    // The last warp is the producer (bulk copies of the next column, metadata two columns
    // ahead: long-latency global loads); when the block has a warp to spare it is kept out
    // of the FFT and block-product phases so that nobody waits for it at their barriers.
    static constexpr int WORK_WARPS = SC_WARPS < NWARPS ? NWARPS - 1 : NWARPS;
    static constexpr int NW_MAX = (NN + WORK_WARPS - 1) / WORK_WARPS;  // nodes per warp (even split)
    // U row stride: even (rows are 16-byte aligned targets of the bulk copies) and with the
    // fewest bank conflicts for the FFT gathers (a half-warp reads (node, a), node < NW_MAX,
    // a < P at doubles node * SU + a + const)
    static constexpr int su_conflicts(int v) {
        int worst = 0;
        for (int b = 0; b < 16; ++b) {
            int c = 0;
            for (int dn = 0; dn < NW_MAX; ++dn)
                for (int a = 0; a < P; ++a) c += ((dn * v + a) & 15) == b;
            worst = c > worst ? c : worst;
        }
        return worst;
    }
    static constexpr int su() {
        int best = NTP, bc = 1 << 30;
        for (int v = NTP; v < NTP + 32; v += 2) {
            const int c = su_conflicts(v);
            if (c < bc) { bc = c; best = v; }
        }
        return best;
    }
    static constexpr int SU = su(), SW = CircCfg<N_, NC_, P_>::SW;
    static constexpr int U_DOUBLES = NN * SU;
    static constexpr unsigned TX_BYTES = (unsigned)(sizeof(double) * (NN * NT + 2 * NN));
    static constexpr int W_DOUBLES = NN * SW;
    static constexpr int T_DOUBLES = M * P * P * 2;            // block spectrum (double2 [M][P][P])
    static constexpr int TAB_DOUBLES = (3 * NT + (NT + 1) / 2 + 1) & ~1;  // wt, wt cos, wt sin + int bits
    static constexpr int COEF_DOUBLES = 2 * NN;
    static constexpr size_t SMEM = sizeof(double) * (size_t)(2 * U_DOUBLES + W_DOUBLES + T_DOUBLES +
                                                             TAB_DOUBLES + 2 * COEF_DOUBLES);
    // Real rows of W (the inverse transform's output, read by the streaming phase): each warp
    // rewrites the region of its own nodes, row stride SWR = P (mod 16) so that the 8-byte
    // stores of a half-warp ((node, a) at doubles node * SWR + a + const) hit distinct banks
    // -- the spectrum layout of the same region needs the even stride SW for 16-byte slots.
    static constexpr int swr() {
        int v = NT;
        while ((v & 15) != (P & 15)) ++v;
        return v <= SW ? v : SW;
    }
    static constexpr int SWR = swr();
    __host__ __device__ static constexpr int node_warp(int n) {      // owner of node n in the FFT phases
        int w = 0;
        while (((w + 1) * NN) / WORK_WARPS <= n) ++w;
        return w;
    }
    __host__ __device__ static constexpr int wrow(int n) {           // offset (doubles) of node n's real row
        const int n0 = (node_warp(n) * NN) / WORK_WARPS;
        return n0 * SW + (n - n0) * SWR;
    }
    static constexpr int NA = (NN + 1) / 2;                    // nodes [0, NA) -> half 0, rest -> half 1
    // resident blocks per SM the register allocation aims at: 2 at the benchmark size (320
    // threads, 96 registers), more for the small blocks of low-order / few-cell columns
    static constexpr int MINB = 640 / THREADS < 2 ? 2 : (640 / THREADS > 8 ? 8 : 640 / THREADS);
    static_assert(NW_MAX * P <= 16, "a warp transforms at most 16 sequences");
};

// ---- bulk asynchronous copies (TMA, non-tensor form) completing on an mbarrier ---- //
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const unsigned count) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(b), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, const unsigned bytes,
                                         unsigned long long* bar) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s),
        "l"(gmem), "r"(bytes), "r"(b)
        : "memory");
}

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

// ---- streaming terms of one angle, one half of the spatial nodes ------------- //
// Same arithmetic as stream_theta_circ (apply_circ.cuh) for columns whose faces are all
// conforming (the host routes columns with mortar faces to k_apply_circ); HALF = 0 owns nodes
// [0, NA), HALF = 1 owns [NA, NN).  Row 0 always belongs to half 0 and row N-1 to
// half 1, so the inflow x-face is handled by exactly one of the two threads.
// FULL: all four terms (the solver's call); the term tests are compiled out.
// Order: neighbour-trace loads are issued first and consumed last (L2 latency hidden
// behind the two derivative passes over shared memory).
// Kernel arguments the streaming phase needs, kept in shared memory: as parameters of the
// out-of-line stream7 they would travel through the stack (more arguments than the ABI has
// registers for), and the streaming phase would wait for those local-memory loads behind its own
// global trace loads (13 % of its stall samples, profiles/r02_q_*).
struct StreamArgs7 {
    const double* x;
    const double* ghost;
    double* y;
    unsigned terms;
    int interior_only;
};

template <int N, int HALF, bool FULL, typename Cfg, bool INNER = false>
__device__ __noinline__ void stream7(const ColMeta& cm, const int th, const int cur, const StreamArgs7& sa) {
    // the block's dynamic shared memory (layout: k_apply_circ7), addressed from its base so that
    // the call needs four argument registers
    extern __shared__ __align__(16) double smem7[];
    const double* __restrict__ us = smem7 + cur * Cfg::U_DOUBLES;
    const double* __restrict__ tab = smem7 + 2 * Cfg::U_DOUBLES + Cfg::W_DOUBLES + Cfg::T_DOUBLES;
    const double* __restrict__ coef_s = tab + Cfg::TAB_DOUBLES + cur * Cfg::COEF_DOUBLES;
    const unsigned terms = sa.terms;
    const double* __restrict__ ws = (FULL || (terms & RTB200_TERM_SCAT)) ? smem7 + 2 * Cfg::U_DOUBLES + th : nullptr;
    const int interior_only = sa.interior_only;
    const double* __restrict__ x = sa.x;
    const double* __restrict__ ghost = sa.ghost;
    __builtin_assume(__isGlobal(x));        // read back from shared memory: tell the compiler the state space
    __builtin_assume(__isGlobal(ghost));
    constexpr int NN = N * N, NT = Cfg::NT, SU = Cfg::SU, SW = Cfg::SW, NTP = Cfg::NTP;
    constexpr int NA = Cfg::NA;
#define RT_MINE(i, j) (HALF == 0 ? ((i) * N + (j) < NA) : ((i) * N + (j) >= NA))
    const double* wq = c_w + N * RT_MAXN;
    const double* S = c_s + N * RT_MAXN * RT_MAXN;
    const bool t_mass = FULL || (terms & RTB200_TERM_MASS), t_ic = FULL || (terms & RTB200_TERM_ICONV);
    const bool t_bc = FULL || (terms & RTB200_TERM_BCONV);
    const bool t_sc = FULL || ws != nullptr;

    const int bits = reinterpret_cast<const int*>(tab + 3 * NT)[th];
    const int Fx = (bits & 1) ? 2 : 0, Fy = (bits & 2) ? 3 : 1;        // the two inflow faces
    const int bdry = INNER ? 0 : cm.bdry;       // INNER: column off the domain boundary (block-uniform)

    // inflow x-face: row N-1 when Fx == 0 (half 1), row 0 when Fx == 2 (half 0)
    constexpr int IX = HALF == 0 ? 0 : N - 1;
    const bool own_x = (HALF == 0) == (Fx == 2);
    const bool in_x = t_bc && own_x && !((bdry >> Fx) & 1);
    const bool in_y = t_bc && !((bdry >> Fy) & 1);
    double tx[N], ty[N];
    {
        const long long base = cm.face_base[2 * Fx];
        const double* __restrict__ src = (base >= 0 ? x + base : ghost + (-(base + 2))) + th;
        __builtin_assume(__isGlobal(src));
        const int rs = cm.rs[2 * Fx];
#pragma unroll
        for (int q = 0; q < N; ++q) tx[q] = in_x ? src[(int64_t)q * rs] : 0.0;
    }
    {
        const long long base = cm.face_base[2 * Fy];
        const double* __restrict__ src = (base >= 0 ? x + base : ghost + (-(base + 2))) + th;
        __builtin_assume(__isGlobal(src));
        const int rs = cm.rs[2 * Fy];
#pragma unroll
        for (int q = 0; q < N; ++q)
            ty[q] = (in_y && (RT_MINE(q, 0) || RT_MINE(q, N - 1))) ? src[(int64_t)q * rs] : 0.0;
    }

    // face weights wt cos(theta - F pi/2) = (ct, st, -ct, -st) for F = 0..3 (the table th_wF
    // holds the same numbers up to the rounding of the shifted cosine)
    const double wt = tab[th], ct = tab[NT + th], st = tab[2 * NT + th];
    const double wF_inx = Fx == 0 ? ct : -ct, wF_iny = Fy == 1 ? st : -st;
    const double hdx = cm.hdx, hdy = cm.hdy;
    const double* __restrict__ ut = us + th;
    double v[N][N];

    // ---- pass Y: rows of u: y-derivative, mass, outflow faces -------------------- //
    {
        const double cout_x = t_bc ? -(hdy * wF_inx) : 0.0;   // the outflow x-face is Fx ^ 2
        const double cout_y = t_bc ? -(hdx * wF_iny) : 0.0;
        const bool out_lo_x = Fx == 0, out_lo_y = Fy == 1;    // outflow through face 2 (i = 0) / face 3 (j = 0)
        const double mw = t_mass ? wt : 0.0;
        const double hst = t_ic ? -(hdx * st) : 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (!(RT_MINE(i, 0) || RT_MINE(i, N - 1))) continue;
            double row[N];
#pragma unroll
            for (int k = 0; k < N; ++k) row[k] = ut[(i * N + k) * SU];
            const double cxi = hst * wq[i];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                if (!RT_MINE(i, j)) continue;
                const int n = i * N + j;
                double s = S[j * RT_MAXN] * row[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[j * RT_MAXN + k], row[k], s);
                double acc = cxi * s;
                acc = fma(coef_s[n] * mw, row[j], acc);
                if (i == 0 && out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
                if (i == N - 1 && !out_lo_x) acc = fma(cout_x * wq[j], row[j], acc);
                if (j == 0 && out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
                if (j == N - 1 && !out_lo_y) acc = fma(cout_y * wq[i], row[j], acc);
                v[i][j] = acc;
            }
        }
    }

    // ---- pass X: v[i][j] -= (dy/2) w_j ct sum_k S[i][k] u[k][j] ------------------ //
    if (t_ic) {
        const double hct = -(hdy * ct);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double col[N];
#pragma unroll
            for (int k = 0; k < N; ++k) col[k] = ut[(k * N + j) * SU];
            const double cyj = hct * wq[j];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (!RT_MINE(i, j)) continue;
                double s = S[i * RT_MAXN] * col[0];
#pragma unroll
                for (int k = 1; k < N; ++k) s = fma(S[i * RT_MAXN + k], col[k], s);
                v[i][j] = fma(cyj, s, v[i][j]);
            }
        }
    }

    // ---- inflow faces (conforming neighbours: raw traces) -------------------------- //
    if (t_bc) {
        const double cin_x = hdy * wF_inx, cin_y = hdx * wF_iny;
        if (own_x) {
#pragma unroll
            for (int q = 0; q < N; ++q) v[IX][q] = fma(cin_x * wq[q], tx[q], v[IX][q]);
        }
        if (Fy == 1) {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (RT_MINE(q, N - 1)) v[q][N - 1] = fma(cin_y * wq[q], ty[q], v[q][N - 1]);
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q)
                if (RT_MINE(q, 0)) v[q][0] = fma(cin_y * wq[q], ty[q], v[q][0]);
        }
    }

    // ---- scattering combine, interior restriction (get_masks.py:80-110), store ----- //
    const bool msk = !INNER && interior_only != 0;
    const bool m0 = msk && (bdry & 1) && (bits & 16), m1 = msk && (bdry & 2) && (bits & 32);
    const bool m2 = msk && (bdry & 4) && (bits & 64), m3 = msk && (bdry & 8) && (bits & 128);
    asm volatile("" ::: "memory");          // y is fetched here, not at the top (it would be spilled across the phase)
    double* __restrict__ y = sa.y;
    __builtin_assume(__isGlobal(y));
    double* yc = y + cm.slab_off + th;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (!RT_MINE(i, j)) continue;
            const int n = i * N + j;
            double acc = v[i][j];
            if (t_sc) acc = fma(-coef_s[NN + n], ws[Cfg::wrow(n)], acc);
            bool masked = false;
            if (i == N - 1) masked = masked || m0;
            if (j == N - 1) masked = masked || m1;
            if (i == 0) masked = masked || m2;
            if (j == 0) masked = masked || m3;
            yc[(int64_t)n * NTP] = masked ? 0.0 : acc;
        }
    }
#undef RT_MINE
}

// ---- scattering, per column: F (FFT) | M (block product) | I (inverse FFT) --------- //
// F / I: a warp owns whole spatial nodes (an even split of the NN nodes over the warps,
//        at most 16 / P each); lanes l and l ^ 16 share one sequence (node, a).
//        F: U -> spectrum slots in W;  I: slots -> real row W[node][theta], in place
//        (only the owning warp touches a node's row: __syncwarp is enough).
// M:     lane = spatial node, the frequency slot k is warp-uniform: the P x P complex
//        block T[k] is read from shared memory with warp-uniform (broadcast) 16-byte
//        loads -- no per-lane table registers (a constant-bank table was measured slower:
//        the uniform loads miss in the constant cache);
//        each lane reads X[k][0..P) of its node once, forms all P outputs (4 P^2 DFMA)
//        and writes them back in place.  Warp w takes slots w, w + NWARPS, ...
template <typename Cfg>
struct CircTab7 {
    double2 v[Cfg::M * Cfg::P * Cfg::P];       // [k][r][a], see rt/tables.py:circulant_spectrum
};

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

// one complex slot: y[r] = sum_q T[k][r][q] u[q], in place
template <typename Cfg>
__device__ __forceinline__ void mult7_one(double2* const w2, const double2* __restrict__ tk, const bool on) {
    constexpr int P = Cfg::P;
    double2 u[P];
#pragma unroll
    for (int q = 0; q < P; ++q) u[q] = w2[q];
#pragma unroll
    for (int r = 0; r < P; ++r) {
        double y0 = tk[r * P].x * u[0].x, y1 = -(tk[r * P].y * u[0].y);
        double y2 = tk[r * P].x * u[0].y, y3 = tk[r * P].y * u[0].x;
#pragma unroll
        for (int q = 1; q < P; ++q) {
            y0 = fma(tk[r * P + q].x, u[q].x, y0);
            y1 = fma(-tk[r * P + q].y, u[q].y, y1);
            y2 = fma(tk[r * P + q].x, u[q].y, y2);
            y3 = fma(tk[r * P + q].y, u[q].x, y3);
        }
        if (on) w2[r] = make_double2(y0 + y1, y2 + y3);
    }
}

template <typename Cfg>
__device__ __noinline__ void mult7(double* __restrict__ W, const double2* __restrict__ Ts, const int warp_u,
                                   const int lane) {
    constexpr int P = Cfg::P, NN = Cfg::NN, SW = Cfg::SW, M = Cfg::M, WW = Cfg::WORK_WARPS;
    static_assert(NN <= 32, "one lane per spatial node");
    if (warp_u >= WW) return;
    const bool on = lane < NN;
    double2* const wn = reinterpret_cast<double2*>(W) + (on ? lane : 0) * (SW / 2);
    int k = warp_u;
    if (k == 0) {                    // the two real bins (X[0], X[NC/2]): no cross terms
        double2* const w2 = wn;
        double2 u[P];
#pragma unroll
        for (int q = 0; q < P; ++q) u[q] = w2[q];
#pragma unroll
        for (int r = 0; r < P; ++r) {
            double yr = Ts[r * P].x * u[0].x, yi = Ts[r * P].y * u[0].y;
#pragma unroll
            for (int q = 1; q < P; ++q) {
                yr = fma(Ts[r * P + q].x, u[q].x, yr);
                yi = fma(Ts[r * P + q].y, u[q].y, yi);
            }
            if (on) w2[r] = make_double2(yr, yi);
        }
        k += WW;
    }
    // pairs of slots (k, k + WW) interleaved: 8 independent chains per output row
#pragma unroll 1
    for (; k + WW < M; k += 2 * WW) {
        double2* const wa = wn + k * P;
        double2* const wb = wn + (k + WW) * P;
        const double2* __restrict__ ta = Ts + k * P * P;
        const double2* __restrict__ tb = Ts + (k + WW) * P * P;
        double2 ua[P], ub[P];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            ua[q] = wa[q];
            ub[q] = wb[q];
        }
#pragma unroll
        for (int r = 0; r < P; ++r) {
            double a0 = ta[r * P].x * ua[0].x, a1 = -(ta[r * P].y * ua[0].y);
            double a2 = ta[r * P].x * ua[0].y, a3 = ta[r * P].y * ua[0].x;
            double b0 = tb[r * P].x * ub[0].x, b1 = -(tb[r * P].y * ub[0].y);
            double b2 = tb[r * P].x * ub[0].y, b3 = tb[r * P].y * ub[0].x;
#pragma unroll
            for (int q = 1; q < P; ++q) {
                a0 = fma(ta[r * P + q].x, ua[q].x, a0);
                a1 = fma(-ta[r * P + q].y, ua[q].y, a1);
                a2 = fma(ta[r * P + q].x, ua[q].y, a2);
                a3 = fma(ta[r * P + q].y, ua[q].x, a3);
                b0 = fma(tb[r * P + q].x, ub[q].x, b0);
                b1 = fma(-tb[r * P + q].y, ub[q].y, b1);
                b2 = fma(tb[r * P + q].x, ub[q].y, b2);
                b3 = fma(tb[r * P + q].y, ub[q].x, b3);
            }
            if (on) {
                wa[r] = make_double2(a0 + a1, a2 + a3);
                wb[r] = make_double2(b0 + b1, b2 + b3);
            }
        }
    }
    if (k < M) mult7_one<Cfg>(wn + k * P, Ts + k * P * P, on);
}

// Same product when every block T[k] is Hermitian (scattering kernels that are symmetric
// in (theta, theta'), e.g. any Phi(|theta - theta'|): G = diag(wt) Phi diag(wt) is then
// symmetric and its block spectrum Hermitian; checked numerically on the host).  Only the
// upper triangle is read: T[r][q] feeds y[r] += T[r][q] u[q] and y[q] += conj(T[r][q]) u[r],
// 15 instead of 25 broadcast loads per slot, all P outputs accumulated at once (4 P
// independent chains, so slots are not paired).
template <typename Cfg>
__device__ __noinline__ void mult7_herm(double* __restrict__ W, const double2* __restrict__ Ts,
                                        const int warp_u, const int lane) {
    constexpr int P = Cfg::P, NN = Cfg::NN, SW = Cfg::SW, M = Cfg::M, WW = Cfg::WORK_WARPS;
    static_assert(NN <= 32, "one lane per spatial node");
    if (warp_u >= WW) return;
    const bool on = lane < NN;
    double2* const wn = reinterpret_cast<double2*>(W) + (on ? lane : 0) * (SW / 2);
#pragma unroll 1
    for (int k = warp_u; k < M; k += WW) {
        double2* const w2 = wn + k * P;
        const double2* __restrict__ tk = Ts + k * P * P;
        double2 u[P];
#pragma unroll
        for (int q = 0; q < P; ++q) u[q] = w2[q];
        if (k == 0) {            // two real symmetric products: re <- T.x u.x, im <- T.y u.y
            double yr[P], yi[P];
#pragma unroll
            for (int r = 0; r < P; ++r) {
                const double2 t = tk[r * P + r];
                yr[r] = t.x * u[r].x;
                yi[r] = t.y * u[r].y;
            }
#pragma unroll
            for (int r = 0; r < P; ++r)
#pragma unroll
                for (int q = r + 1; q < P; ++q) {
                    const double2 t = tk[r * P + q];
                    yr[r] = fma(t.x, u[q].x, yr[r]);
                    yi[r] = fma(t.y, u[q].y, yi[r]);
                    yr[q] = fma(t.x, u[r].x, yr[q]);
                    yi[q] = fma(t.y, u[r].y, yi[q]);
                }
            if (on) {
#pragma unroll
                for (int r = 0; r < P; ++r) w2[r] = make_double2(yr[r], yi[r]);
            }
        } else {
            double y0[P], y1[P], y2[P], y3[P];       // re = y0 - y1, im = y2 + y3
#pragma unroll
            for (int r = 0; r < P; ++r) {
                const double2 t = tk[r * P + r];
                y0[r] = t.x * u[r].x;
                y1[r] = t.y * u[r].y;
                y2[r] = t.x * u[r].y;
                y3[r] = t.y * u[r].x;
            }
#pragma unroll
            for (int r = 0; r < P; ++r)
#pragma unroll
                for (int q = r + 1; q < P; ++q) {
                    const double2 t = tk[r * P + q];
                    // y[r] += t u[q]
                    y0[r] = fma(t.x, u[q].x, y0[r]);
                    y1[r] = fma(t.y, u[q].y, y1[r]);
                    y2[r] = fma(t.x, u[q].y, y2[r]);
                    y3[r] = fma(t.y, u[q].x, y3[r]);
                    // y[q] += conj(t) u[r]
                    y0[q] = fma(t.x, u[r].x, y0[q]);
                    y1[q] = fma(-t.y, u[r].y, y1[q]);
                    y2[q] = fma(t.x, u[r].y, y2[q]);
                    y3[q] = fma(-t.y, u[r].x, y3[q]);
                }
            if (on) {
#pragma unroll
                for (int r = 0; r < P; ++r) w2[r] = make_double2(y0[r] - y1[r], y2[r] + y3[r]);
            }
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


// ---- scattering, E/O form (default): no cross-lane exchange at all ------------------ //
// The two lanes of a sequence transform the even / the odd samples (length NC/2, as above)
// and store their HALF-LENGTH spectra E, O as they are; the last radix-2 stage of the
// forward transform and the first of the inverse move into the block-product phase, whose
// thread (lane = node, task k warp-uniform) owns the slots k and NC/2 - k together:
//     2 X[k]        = 2 E[k] + w^k 2 O[k],        2 X[NC/2 - k] = conj(2 E[k] - w^k 2 O[k])
//     Y[k] = T[k] 2 X[k],   Y[NC/2 - k] = T[NC/2 - k] 2 X[NC/2 - k]
//     A[k] = Y[k] + conj(Y[NC/2 - k])             -> even samples of the result
//     B[k] = (Y[k] - conj(Y[NC/2 - k])) w^{-k}    -> odd samples
// (w = exp(-2 pi i / NC); task 0 owns the packed real bins: slots 0 and NC/4).  Against the
// shuffle form (split_forward / split_inverse): no SHFL, no per-lane role selects, the
// same FP64 work, NC/4 block-product tasks of two slots each -- one round over the
// work warps at the benchmark size instead of two uneven ones.
// Spectrum layout per node (double2 units): [(2 j + t) P + a], j < NC/4: entry j of the
// even (t = 0) / odd (t = 1) half-spectrum of sequence a; entry 0 packs the two real bins.
template <typename Cfg>
__device__ __noinline__ void fft_f8(const double* __restrict__ U, double* __restrict__ W,
                                    const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, Q = Cfg::Q, SU = Cfg::SU, SW = Cfg::SW;
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
    double hr[Q], hi[Q];
    rtfft::Real<NC / 2>::forward(xs, hr, hi);
    if (seq_on) {
        double2* const w = reinterpret_cast<double2*>(W + node * SW) + t * P + a;
#pragma unroll
        for (int j = 0; j < Q; ++j) w[2 * j * P] = make_double2(hr[j], hi[j]);
    }
}

// y[r] = sum_q T[r][q] u[q] (complex); HERM: only the upper triangle of T is read
template <int P, bool HERM>
__device__ __forceinline__ void cprod8(const double2* __restrict__ tk, const double2 (&u)[P], double2 (&y)[P]) {
    if constexpr (HERM) {
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const double2 t = tk[r * P + r];
            y[r].x = fma(-t.y, u[r].y, t.x * u[r].x);
            y[r].y = fma(t.y, u[r].x, t.x * u[r].y);
        }
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int q = r + 1; q < P; ++q) {
                const double2 t = tk[r * P + q];
                y[r].x = fma(t.x, u[q].x, y[r].x);          // y[r] += t u[q]
                y[r].y = fma(t.x, u[q].y, y[r].y);
                y[q].x = fma(t.x, u[r].x, y[q].x);          // y[q] += conj(t) u[r]
                y[q].y = fma(t.x, u[r].y, y[q].y);
                y[r].x = fma(-t.y, u[q].y, y[r].x);
                y[r].y = fma(t.y, u[q].x, y[r].y);
                y[q].x = fma(t.y, u[r].y, y[q].x);
                y[q].y = fma(-t.y, u[r].x, y[q].y);
                if (q == P - 1 && (r & 1)) asm volatile("" ::: "memory");   // bounds the table loads in flight
            }
    } else {
#pragma unroll
        for (int r = 0; r < P; ++r) {
            double re = tk[r * P].x * u[0].x, im = tk[r * P].x * u[0].y;
            double r2 = -(tk[r * P].y * u[0].y), i2 = tk[r * P].y * u[0].x;
#pragma unroll
            for (int q = 1; q < P; ++q) {
                const double2 t = tk[r * P + q];
                re = fma(t.x, u[q].x, re);
                im = fma(t.x, u[q].y, im);
                r2 = fma(-t.y, u[q].y, r2);
                i2 = fma(t.y, u[q].x, i2);
            }
            y[r] = make_double2(re + r2, im + i2);
            asm volatile("" ::: "memory");      // one row of table loads in flight at a time
        }
    }
}

// two real products at once: y.x = sum_q T.x u.x, y.y = sum_q T.y u.y (the packed real bins)
template <int P, bool HERM>
__device__ __forceinline__ void rprod8(const double2* __restrict__ tk, const double2 (&u)[P], double2 (&y)[P]) {
    if constexpr (HERM) {
#pragma unroll
        for (int r = 0; r < P; ++r) {
            const double2 t = tk[r * P + r];
            y[r] = make_double2(t.x * u[r].x, t.y * u[r].y);
        }
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int q = r + 1; q < P; ++q) {
                const double2 t = tk[r * P + q];
                y[r].x = fma(t.x, u[q].x, y[r].x);
                y[r].y = fma(t.y, u[q].y, y[r].y);
                y[q].x = fma(t.x, u[r].x, y[q].x);
                y[q].y = fma(t.y, u[r].y, y[q].y);
            }
    } else {
#pragma unroll
        for (int r = 0; r < P; ++r) {
            double yr = tk[r * P].x * u[0].x, yi = tk[r * P].y * u[0].y;
#pragma unroll
            for (int q = 1; q < P; ++q) {
                yr = fma(tk[r * P + q].x, u[q].x, yr);
                yi = fma(tk[r * P + q].y, u[q].y, yi);
            }
            y[r] = make_double2(yr, yi);
            asm volatile("" ::: "memory");
        }
    }
}

template <typename Cfg, bool HERM>
__device__ __noinline__ void mult8(double* __restrict__ W, const double2* __restrict__ Ts, const int warp_u,
                                   const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, NN = Cfg::NN, SW = Cfg::SW, H = Cfg::H, Q = Cfg::Q;
    constexpr int WW = Cfg::WORK_WARPS;
    static_assert(NN <= 32, "one lane per spatial node");
    if (warp_u >= WW) return;
    const bool on = lane < NN;
    double2* const wn = reinterpret_cast<double2*>(W) + (on ? lane : 0) * (SW / 2);
#pragma unroll 1
    for (int k = warp_u; k < Q; k += WW) {                      // warp-uniform
        double2* const we = wn + 2 * k * P;                     // even entries [a]; odd entries at + P
        double2 ua[P], ub[P];
        if (k == 0) {
            // slot 0 = (X[0], X[NC/2]) = (E0 + O0, E0 - O0); slot NC/4 = 2 X[NC/4] = (2 E[NC/4], -2 O[NC/4])
#pragma unroll
            for (int a = 0; a < P; ++a) {
                const double2 e = we[a], o = we[P + a];
                ua[a] = make_double2(e.x + o.x, e.x - o.x);
                ub[a] = make_double2(2.0 * e.y, -2.0 * o.y);
            }
            double2 y0[P], yq[P];
            rprod8<P, HERM>(Ts, ua, y0);
            asm volatile("" ::: "memory");
            cprod8<P, HERM>(Ts + Q * P * P, ub, yq);
            if (on) {
#pragma unroll
                for (int a = 0; a < P; ++a) {
                    we[a] = make_double2(y0[a].x + y0[a].y, 2.0 * yq[a].x);
                    we[P + a] = make_double2(y0[a].x - y0[a].y, -2.0 * yq[a].y);
                }
            }
        } else {
            const double c = c_fft_cos[k * (64 / NC)], s = c_fft_sin[k * (64 / NC)];   // w^k = c - i s
#pragma unroll
            for (int a = 0; a < P; ++a) {
                const double2 e = we[a], o = we[P + a];
                const double wr = fma(c, o.x, s * o.y), wi = fma(c, o.y, -(s * o.x));
                ua[a] = make_double2(e.x + wr, e.y + wi);
                ub[a] = make_double2(e.x - wr, wi - e.y);
            }
            double2 ya[P];
            cprod8<P, HERM>(Ts + k * P * P, ua, ya);
            asm volatile("" ::: "memory");      // keeps the second block's table loads (and their registers) behind the first product
            double2 yb[P];
            cprod8<P, HERM>(Ts + (H - k) * P * P, ub, yb);
#pragma unroll
            for (int r = 0; r < P; ++r) {
                const double dr = ya[r].x - yb[r].x, di = ya[r].y + yb[r].y;
                if (on) {
                    we[r] = make_double2(ya[r].x + yb[r].x, ya[r].y - yb[r].y);
                    we[P + r] = make_double2(fma(dr, c, -(di * s)), fma(dr, s, di * c));
                }
            }
        }
    }
}

template <typename Cfg>
__device__ __noinline__ void fft_i8(double* __restrict__ W, const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, Q = Cfg::Q, SW = Cfg::SW;
    constexpr int NN = Cfg::NN;
    if (warp >= Cfg::WORK_WARPS) return;                        // warp-uniform
    const int n0 = (warp * NN) / Cfg::WORK_WARPS;
    const int nw = ((warp + 1) * NN) / Cfg::WORK_WARPS - n0;
    if (nw == 0) return;
    const int t = lane >> 4, sq = lane & 15;
    const bool seq_on = sq < nw * P;
    const int dn = seq_on ? sq / P : 0, a = seq_on ? sq - dn * P : 0;
    const int node = n0 + dn;
    const double2* const w = reinterpret_cast<const double2*>(W + node * SW) + t * P + a;
    double hr[Q], hi[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const double2 v = w[2 * j * P];
        hr[j] = v.x;
        hi[j] = v.y;
    }
    __syncwarp();                                               // the real rows overwrite the spectrum
    double xs[H];
    rtfft::Real<NC / 2>::inverse(hr, hi, xs);
    if (seq_on) {
        double* dst = W + n0 * SW + dn * Cfg::SWR + a + t * P;      // real-row layout, see Circ7Cfg::wrow
#pragma unroll
        for (int mm = 0; mm < H; ++mm) dst[2 * mm * P] = xs[mm];
    }
}

// barrier among the work warps (all warps when the block has no spare producer warp)
template <typename Cfg>
__device__ __forceinline__ void work_sync(const int warp) {
    if constexpr (Cfg::WORK_WARPS < Cfg::NWARPS) {
        if (warp < Cfg::WORK_WARPS) asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::WORK_WARPS * 32) : "memory");
    } else {
        __syncthreads();
    }
}

template <int N, int NC, int P>
__global__ void __launch_bounds__(Circ7Cfg<N, NC, P>::THREADS, Circ7Cfg<N, NC, P>::MINB)
k_apply_circ7(DevMesh m, const int32_t* __restrict__ cols, const int ncols, int sg, unsigned terms,
              int interior_only, int herm, const double* __restrict__ x, const double* __restrict__ ghost,
              double* __restrict__ y, const __grid_constant__ CircTab7<Circ7Cfg<N, NC, P>> T,
              const int n_free, const unsigned* __restrict__ halo_flag, const unsigned halo_epoch,
              const int* __restrict__ resume_in, int* __restrict__ resume_out, const int mode) {
    using Cfg = Circ7Cfg<N, NC, P>;
    const bool eo = mode & 1, inner_fast = (mode & 2) != 0;
    constexpr int NN = Cfg::NN, NT = Cfg::NT, NTP = Cfg::NTP;
    constexpr int SU = Cfg::SU, THREADS = Cfg::THREADS;
    extern __shared__ __align__(16) double smem7[];
    double* const Ub = smem7;                                   // [2][U_DOUBLES]
    double* const W = smem7 + 2 * Cfg::U_DOUBLES;               // [NN][SW]
    double* const Ts = W + Cfg::W_DOUBLES;                      // block spectrum T[k][r][a]
    double* const tab = Ts + Cfg::T_DOUBLES;                    // per-angle tables
    double* const coefb = tab + Cfg::TAB_DOUBLES;               // [2][2 NN]
    __shared__ ColMeta7 s_meta[3];
    __shared__ __align__(8) unsigned long long s_bar[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stride = gridDim.x;
    // Multi-GPU (rt_apply_with_halo): the columns from position n_free on read ghost traces.
    // Launch 1 runs while the exchange is in flight: a block that reaches its first such column
    // before the exchange has published its epoch in *halo_flag does NOT wait (a persistent
    // grid that spins can keep the exchange's own kernels off the SMs for ever): it notes
    // where it stopped in resume_out and exits; launch 2, ordered after the exchange, starts
    // every block at resume_in.  Blocks that find the flag raised simply carry on.
    int it = resume_in ? resume_in[blockIdx.x] : blockIdx.x;
    if (it >= ncols) {
        if (resume_out && threadIdx.x == 0) resume_out[blockIdx.x] = it;
        return;
    }
    __shared__ int s_go;
    __shared__ StreamArgs7 s_args;
    if (threadIdx.x == 0) s_args = StreamArgs7{x, ghost, y, terms, interior_only};   // published by the prologue barrier

    // ---- prologue: per-angle tables, block spectrum, metadata of the first two columns -- //
    {
        const int64_t tp0 = m.sig_th0[sg];
        for (int e = tid; e < NT; e += THREADS) {
            tab[e] = m.th_wt[tp0 + e];
            tab[NT + e] = m.th_ct[tp0 + e];
            tab[2 * NT + e] = m.th_st[tp0 + e];
            reinterpret_cast<int*>(tab + 3 * NT)[e] = m.th_bits[tp0 + e];
        }
        for (int e = tid; e < Cfg::M * Cfg::P * Cfg::P; e += THREADS)
            reinterpret_cast<double2*>(Ts)[e] = T.v[e];
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0 && lane < 9) load_col_meta7(m, cols[it], s_meta[0], lane);
    if (warp == Cfg::NWARPS - 1 && lane < 9 && it + stride < ncols)
        load_col_meta7(m, cols[it + stride], s_meta[1], lane);
    __syncthreads();

    // One lane issues the bulk copies of a column (NN rows of the slab + the coefficient
    // samples); they complete on the buffer's mbarrier.
    auto prefetch = [&](const int buf, const ColMeta7& cm) {
        if (warp == Cfg::NWARPS - 1 && lane == 16) {
            // order the block's earlier generic-proxy reads of the buffer before the
            // async-proxy writes (the block barrier made them visible to this thread)
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            mbar_expect_tx(&s_bar[buf], Cfg::TX_BYTES);
            const double* __restrict__ src = x + cm.cm.slab_off;
            double* dst = Ub + buf * Cfg::U_DOUBLES;
#pragma unroll 1
            for (int node = 0; node < NN; ++node)
                bulk_g2s(dst + node * SU, src + node * NTP, (unsigned)(sizeof(double) * NT), &s_bar[buf]);
            bulk_g2s(coefb + buf * Cfg::COEF_DOUBLES, m.coef + cm.coef_off,
                     (unsigned)(sizeof(double) * 2 * NN), &s_bar[buf]);
        }
    };
    prefetch(0, s_meta[0]);

    const bool full = terms == RTB200_TERM_ALL;
    const int half = tid >= Cfg::NTW ? 1 : 0;
    const int th = tid - half * Cfg::NTW;
    int cur = 0, ms = 0;
    int n_wait = n_free;     // first position that still has to check the halo flag (ncols once it was seen)
    unsigned phase = 0;      // bit b: parity of the next completion of buffer b
    for (; it < ncols; it += stride) {
        if (it >= n_wait) {
            if (tid == 0) {
                unsigned seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(halo_flag) : "memory");
                s_go = (int)(seen - halo_epoch) >= 0;
            }
            __syncthreads();
            if (!s_go) {
                mbar_wait(&s_bar[cur], (phase >> cur) & 1u);     // the copy of this column is in flight
                break;
            }
            n_wait = ncols;
        }
        __syncthreads();     // everybody is done with the previous column (U[cur ^ 1], W are free)
        const int ms1 = ms == 2 ? 0 : ms + 1, ms2 = ms1 == 2 ? 0 : ms1 + 1;
        if (it + stride < ncols) prefetch(cur ^ 1, s_meta[ms1]);
        if (warp == Cfg::NWARPS - 1 && lane < 9 && it + 2 * stride < ncols)
            load_col_meta7(m, cols[it + 2 * stride], s_meta[ms2], lane);

        mbar_wait(&s_bar[cur], (phase >> cur) & 1u);     // U[cur] / coef[cur] landed
        phase ^= 1u << cur;
        const double* U = Ub + cur * Cfg::U_DOUBLES;
        if (terms & RTB200_TERM_SCAT) {
            if (eo) {            // E/O form: the radix-2 stage across the lane pair lives in the block product
                // The barriers between F, M and I only concern the work warps: a named barrier keeps
                // the producer warp (bulk-copy issue, metadata loads from global memory) out of them,
                // so that nobody waits for its latencies before the streaming phase.
                fft_f8<Cfg>(U, W, warp, lane);
                work_sync<Cfg>(warp);
                if (herm)
                    mult8<Cfg, true>(W, reinterpret_cast<const double2*>(Ts), __shfl_sync(0xffffffffu, warp, 0), lane);
                else
                    mult8<Cfg, false>(W, reinterpret_cast<const double2*>(Ts), __shfl_sync(0xffffffffu, warp, 0), lane);
                work_sync<Cfg>(warp);
                fft_i8<Cfg>(W, warp, lane);
            } else {             // shuffle form (RTB200_CIRC7_EO=0)
                fft_f7<Cfg>(U, W, warp, lane);
                __syncthreads();
                if (herm)
                    mult7_herm<Cfg>(W, reinterpret_cast<const double2*>(Ts), __shfl_sync(0xffffffffu, warp, 0), lane);
                else
                    mult7<Cfg>(W, reinterpret_cast<const double2*>(Ts), __shfl_sync(0xffffffffu, warp, 0), lane);
                __syncthreads();
                fft_i7<Cfg>(W, warp, lane);
            }
            __syncthreads();
        }

        // ---- phase S: streaming terms + scattering combine -> y -------------------- //
        {
            const ColMeta& cm = s_meta[ms].cm;
            if (th < NT && tid < 2 * Cfg::NTW) {
                if (full && inner_fast && cm.bdry == 0) {
                    // no face on the domain boundary: every inflow face has a neighbour, no row is masked
                    if (half == 0)
                        stream7<N, 0, true, Cfg, true>(cm, th, cur, s_args);
                    else
                        stream7<N, 1, true, Cfg, true>(cm, th, cur, s_args);
                } else if (full) {
                    if (half == 0)
                        stream7<N, 0, true, Cfg>(cm, th, cur, s_args);
                    else
                        stream7<N, 1, true, Cfg>(cm, th, cur, s_args);
                } else {
                    if (half == 0)
                        stream7<N, 0, false, Cfg>(cm, th, cur, s_args);
                    else
                        stream7<N, 1, false, Cfg>(cm, th, cur, s_args);
                }
            } else if (half == 0 && th < NTP) {             // pad entry of an odd row length
                double* yc = y + cm.slab_off + th;
                for (int n = 0; n < NN; ++n) yc[(int64_t)n * NTP] = 0.0;
            }
        }
        cur ^= 1;
        ms = ms1;
    }
    if (resume_out && tid == 0) resume_out[blockIdx.x] = it;
}
