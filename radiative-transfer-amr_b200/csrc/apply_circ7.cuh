// Persistent, prefetching apply kernel for columns with a UNIFORM angular mesh
// (NC >= 8 equal cells of P nodes; block-circulant scattering, see apply_circ.cuh
// for the algorithm and the table conventions) whose faces are all conforming
// (columns with a mortar face are routed to k_apply_circ by the host).
//
// What changed against k_apply_circ (one block per column, phases in lock step):
//   * persistent blocks (as many as the SMs hold: 2 per SM at the benchmark size) with a DYNAMIC
//     column schedule: the first column of a block is its index, every further one is claimed from
//     a global counter two columns ahead of the one in flight, and the next column's slab +
//     coefficient samples are PREFETCHED with bulk asynchronous copies (cp.async.bulk, one per
//     slab row, completing on an mbarrier) into a second buffer while the current column is
//     processed: the HBM stream and the FP64 phases overlap by construction, and a block that
//     starts late or runs slower simply takes fewer columns;
//   * scattering as F (half-length real FFTs) | M (block product) | I (inverse FFTs): the two
//     lanes (l, l ^ 16) of a sequence (node, a) transform its even / its odd samples (length NC/2,
//     in registers) and store the two half-spectra; the block-product thread (lane = node, task k
//     warp-uniform) owns the slots k and NC/2 - k together and does the last radix-2 stage of the
//     forward and the first of the inverse transform around its two products -- no shuffles;
//     the sequences are packed 16 to a warp across node boundaries (8 transform warps at the
//     benchmark size, two per scheduler), NC/4 product tasks = one per transform warp;
//   * F | M | I are separated by a named barrier over the transform warps only; the last warp is
//     the producer (bulk copies, schedule claims, column metadata two columns ahead in a 3-slot
//     ring) and meets the block at the two full barriers of a column;
//   * streaming terms of an angle split over two threads (first / second half of the spatial
//     nodes), specialised for columns off the domain boundary;
//   * per-angle tables (weights, wt cos, wt sin, upwind bits) and the block spectrum
//     are staged in shared memory once per block.
//
// Shared memory per block at (N, NC, P) = (5, 32, 5): 2 x 32.8 KB (U) + 34.0 KB (W)
// + 6.4 KB (block spectrum) + 4.5 KB (angle tables) + 0.8 KB = 111.3 KB -> 2 blocks / SM.
// Measurements and what bounds the kernel: DESIGN.md 4.1, profiles/r02_summary.md.
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
    // (WORK_WARPS, NW_MAX, node_warp, wrow: whole-node split of the shuffle-form transforms, still
    //  used by the node-wise preconditioner k_pc_circ in precond_circ.cuh)
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
    // E/O transforms (k_apply_circ7): the NN * P sequences are packed 16 to a warp across node
    // boundaries (two lanes each), so the transform phases run on W8 full warps -- 8 at the
    // benchmark size, two per scheduler, where an even split of whole nodes took 9 (3 : 2 : 2 : 2).
    static constexpr int NSEQ = NN * P;
    static constexpr int W8 = (NSEQ + 15) / 16;
    static_assert(W8 <= NWARPS, "transform warps");
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
            if (t_sc) acc = fma(-coef_s[NN + n], ws[n * Cfg::SWR], acc);
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

// ---- scattering, per column: F (half-length FFTs) | M (block product) | I (inverse FFTs) ---- //
// F / I: the NN * P sequences (node, a) are packed 16 to a warp across node boundaries (W8
//        transform warps); lanes l and l ^ 16 share one sequence: one transforms its even, the
//        other its odd samples.  F: U -> half-spectra in W;  I: -> real row W[node * SWR + theta],
//        over the spectra (all loads, barrier among the transform warps, all stores).
// M:     lane = spatial node, the task k (slots k and NC/2 - k) is warp-uniform: the P x P complex
//        blocks T[k] are read from shared memory with warp-uniform (broadcast) 16-byte loads -- no
//        per-lane table registers (a constant-bank table was measured slower).
template <typename Cfg>
struct CircTab7 {
    double2 v[Cfg::M * Cfg::P * Cfg::P];       // [k][r][a], see rt/tables.py:circulant_spectrum
};

// ---- E/O form of the transforms: no cross-lane exchange at all ----------------------- //
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
// same FP64 work, NC/4 block-product tasks of two slots each -- one per transform warp at the
// benchmark size instead of two uneven rounds of single-slot tasks.
// Spectrum layout per node (double2 units): [(2 j + t) P + a], j < NC/4: entry j of the
// even (t = 0) / odd (t = 1) half-spectrum of sequence a; entry 0 packs the two real bins.
template <typename Cfg>
__device__ __noinline__ void fft_f8(const double* __restrict__ U, double* __restrict__ W,
                                    const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, Q = Cfg::Q, SU = Cfg::SU, SW = Cfg::SW;
    if (warp >= Cfg::W8) return;                                // warp-uniform
    const int t = lane >> 4, seq = warp * 16 + (lane & 15);
    const bool seq_on = seq < Cfg::NSEQ;
    const int node = seq_on ? seq / P : 0, a = seq_on ? seq - node * P : 0;
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
    constexpr int WW = Cfg::W8;
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

// barrier among the transform warps (all warps when the block has no spare warp)
template <typename Cfg>
__device__ __forceinline__ void work_sync(const int warp) {
    if constexpr (Cfg::W8 < Cfg::NWARPS) {
        if (warp < Cfg::W8) asm volatile("bar.sync 1, %0;\n" ::"n"(Cfg::W8 * 32) : "memory");
    } else {
        __syncthreads();
    }
}

// Inverse half-length transforms; the real rows W[node * SWR + theta] overwrite the spectra, and a
// node's sequences may sit in two warps: all loads, barrier among the transform warps, all stores.
template <typename Cfg>
__device__ __noinline__ void fft_i8(double* __restrict__ W, const int warp, const int lane) {
    constexpr int NC = Cfg::NC, P = Cfg::P, H = Cfg::H, Q = Cfg::Q, SW = Cfg::SW;
    if (warp >= Cfg::W8) return;                                // warp-uniform
    const int t = lane >> 4, seq = warp * 16 + (lane & 15);
    const bool seq_on = seq < Cfg::NSEQ;
    const int node = seq_on ? seq / P : 0, a = seq_on ? seq - node * P : 0;
    const double2* const w = reinterpret_cast<const double2*>(W + node * SW) + t * P + a;
    double hr[Q], hi[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const double2 v = w[2 * j * P];
        hr[j] = v.x;
        hi[j] = v.y;
    }
    work_sync<Cfg>(warp);                                       // every spectrum entry has been read
    double xs[H];
    rtfft::Real<NC / 2>::inverse(hr, hi, xs);
    if (seq_on) {
        double* dst = W + node * Cfg::SWR + a + t * P;              // uniform real-row stride SWR = P (mod 16)
#pragma unroll
        for (int mm = 0; mm < H; ++mm) dst[2 * mm * P] = xs[mm];
    }
}

template <int N, int NC, int P>
__global__ void __launch_bounds__(Circ7Cfg<N, NC, P>::THREADS, Circ7Cfg<N, NC, P>::MINB)
k_apply_circ7(DevMesh m, const int32_t* __restrict__ cols, const int ncols, int sg, unsigned terms,
              int interior_only, int herm, const double* __restrict__ x, const double* __restrict__ ghost,
              double* __restrict__ y, const __grid_constant__ CircTab7<Circ7Cfg<N, NC, P>> T,
              const int n_free, const unsigned* __restrict__ halo_flag, const unsigned halo_epoch,
              const int* __restrict__ resume_in, int* __restrict__ resume_out, const int mode,
              unsigned* __restrict__ sched) {
    using Cfg = Circ7Cfg<N, NC, P>;
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
    // DYNAMIC COLUMN SCHEDULE.  Position blockIdx.x of the column list is the block's first column;
    // every further position is claimed from the counter sched[0] (position = gridDim.x + old
    // value), two columns ahead of the one being processed -- the depth of the prefetch pipeline.
    // A block that becomes resident late (the halo exchange's own kernels hold SMs for ~35 us at
    // the start of a multi-GPU apply) or runs slower simply claims fewer columns; with static
    // strides it finished that much later and the whole apply with it.  Positions are handed out
    // in increasing order, so the ghost-reading columns at the end of the list stay last.
    // sched[1] counts finished blocks: the last one resets the counters for the next launch
    // (mode bit 3 = first launch of a split apply: the column counter is kept for the second).
    //
    // Multi-GPU (rt_apply_with_halo): the columns from position n_free on read ghost traces.
    // Launch 1 runs while the exchange is in flight: a block that reaches its first such column
    // before the exchange has published its epoch in *halo_flag does NOT wait (a persistent
    // grid that spins can keep the exchange's own kernels off the SMs for ever): it notes its
    // two claimed positions in resume_out and exits; launch 2, ordered after the exchange,
    // starts every block with them and goes on claiming.  Blocks that find the flag raised
    // simply carry on.
    constexpr int NONE = 0x7fffffff;
    __shared__ int s_pos[3];                                    // positions of the columns in s_meta[]
    __shared__ int s_halo_ok;                                   // the halo flag has been seen up
    __shared__ int s_exhausted, s_stopped;                      // a claim came back empty (claimer only) / left with the flag down
    const bool claimer = tid == (Cfg::NWARPS - 1) * 32;         // lane 0 of the producer warp
    auto claim = [&]() -> int {
        const unsigned v = atomicAdd(&sched[0], 1u);
        const long long p = (long long)gridDim.x + (long long)v;
        return p < (long long)ncols ? (int)p : NONE;
    };
    if (claimer) {
        // the current and the next column; the one after next is claimed inside the loop.  A resumed
        // block (launch 2) starts with the two positions it had claimed when it stopped.
        bool exhausted = false;
#pragma unroll 1
        for (int k = 0; k < 2; ++k) {
            int p = NONE;
            if (resume_in) p = resume_in[2 * blockIdx.x + k];
            else if (k == 0) p = (int)blockIdx.x < ncols ? (int)blockIdx.x : NONE;
            if (p == NONE && !exhausted && (resume_in || k > 0)) {
                p = claim();
                exhausted = p == NONE;
            }
            s_pos[k] = p;
        }
        s_pos[2] = NONE;
        s_exhausted = exhausted;
        s_stopped = 0;
        s_halo_ok = 0;
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
    __syncthreads();                                            // s_pos published
    if (s_pos[0] < ncols) {
        if (warp == 0 && lane < 9) load_col_meta7(m, cols[s_pos[0]], s_meta[0], lane);
        if (warp == 1 && lane < 9 && s_pos[1] < ncols) load_col_meta7(m, cols[s_pos[1]], s_meta[1], lane);
    }
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
    if (s_pos[0] < ncols) prefetch(0, s_meta[0]);

    const bool full = terms == RTB200_TERM_ALL;
    const int half = tid >= Cfg::NTW ? 1 : 0;
    const int th = tid - half * Cfg::NTW;
    // Loop state = the iteration count alone (it must live across the out-of-line phase calls):
    //   U buffer cur = iter & 1, metadata slot ms = iter % 3,
    //   mbarrier parity of buffer cur = (iter >> 1) & 1 (its (iter >> 1)-th completion)
    // The count is kept per warp in shared memory and re-read after the phase calls: as a register
    // it is spilled across them, and the local-memory reload at the end of the streaming phase
    // queues behind that phase's global stores.
    __shared__ int s_iter[Cfg::NWARPS];
    if (lane == 0) s_iter[warp] = 0;
    __syncwarp();
    while (true) {
        const int iter = *reinterpret_cast<volatile int*>(&s_iter[warp]);
        const int cur = iter & 1, ms = iter % 3;
        const unsigned par = (unsigned)(iter >> 1) & 1u;
        const int it = s_pos[ms];
        if (it >= ncols) break;
        if (it >= n_free && !s_halo_ok) {
            if (tid == 0) {
                unsigned seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(halo_flag) : "memory");
                s_go = (int)(seen - halo_epoch) >= 0;
            }
            __syncthreads();
            if (!s_go) {
                mbar_wait(&s_bar[cur], par);                     // the copy of this column is in flight
                if (tid == 0) {
                    s_stopped = 1;
                    atomicAdd(&sched[2], 1u);                    // statistics (rtb200_apply_path)
                }
                break;
            }
            if (tid == 0) s_halo_ok = 1;                         // published by the barrier below
        }
        __syncthreads();     // everybody is done with the previous column (U[cur ^ 1], W are free)
        const int ms1 = ms == 2 ? 0 : ms + 1, ms2 = ms1 == 2 ? 0 : ms1 + 1;
        const int it1 = s_pos[ms1];
        if (it1 < ncols) prefetch(cur ^ 1, s_meta[ms1]);
        if (warp == Cfg::NWARPS - 1) {
            // slot ms2 held the previous column (everybody left it at the barrier above): the column
            // after next is claimed now, its metadata has two iterations to arrive
            int p = NONE;
            if (claimer && !s_exhausted && it1 < ncols) {
                p = claim();
                s_exhausted = p == NONE;
            }
            p = __shfl_sync(0xffffffffu, p, 0);
            if (lane == 0) s_pos[ms2] = p;
            if (lane < 9 && p < ncols) load_col_meta7(m, cols[p], s_meta[ms2], lane);
        }

        mbar_wait(&s_bar[cur], par);                     // U[cur] / coef[cur] landed
        const double* U = Ub + cur * Cfg::U_DOUBLES;
        if (terms & RTB200_TERM_SCAT) {
            // The barriers between F, M and I only concern the transform warps: a named barrier keeps
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
            __syncthreads();
        }

        // ---- phase S: streaming terms + scattering combine -> y -------------------- //
        {
            const int iter_s = *reinterpret_cast<volatile int*>(&s_iter[warp]);
            const int cur = iter_s & 1, ms = iter_s % 3;
            const ColMeta& cm = s_meta[ms].cm;
            if (th < NT && tid < 2 * Cfg::NTW) {
                if (full && cm.bdry == 0) {
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
        __syncwarp();
        if (lane == 0) s_iter[warp] = *reinterpret_cast<volatile int*>(&s_iter[warp]) + 1;
        __syncwarp();
    }
    if (tid == 0) {
        const int ms = s_iter[0] % 3;
        if (resume_out) {
#pragma unroll 1
            for (int k = 0; k < 2; ++k) {
                const int slot = (ms + k) % 3;
                resume_out[2 * blockIdx.x + k] = s_stopped ? s_pos[slot] : NONE;
            }
        }
        __threadfence();
        const unsigned d = atomicAdd(&sched[1], 1u);
        if (d == gridDim.x - 1) {                                // last block out: counters ready for the next launch
            sched[1] = 0u;
            if (!(mode & 8)) sched[0] = 0u;
            __threadfence();
        }
    }
}
