// Transport-aware preconditioners for the Krylov solve (sm_100a):
//
//   RTB200_PC_SWEEP    z = mask . T~^{-1} r,   T~ = M~ + BC - IC
//   RTB200_PC_ELEMENT  z = mask . blockdiag(T~)^{-1} r   (element-wise block-Jacobi)
//
// T = M + BC - IC is the streaming part of the reference's system matrix
// (solve.py:52: A = (BC - IC) + (M - S); mass_matrix.py:145,169,
// interior_convection_matrix.py:179-237, boundary_convection_matrix.py:284-430).
// The reference preconditions with PETSc's bjacobi -> ILU(0) (solve.py:93-95), an
// incomplete factorisation of the whole rank-local matrix; its arithmetic lives in PETSc.
// What makes such a factorisation work on this operator is that the upwind face fluxes make
// T block LOWER TRIANGULAR once the elements are ordered along the flow: for the angles of
// one quadrant, the two inflow faces of a column only read columns further upwind.  This
// file uses that structure directly:
//
//   * elements = (column, quadrant) pairs; the host orders them topologically (Kahn) into
//     LEVELS of mutually independent elements -- anti-diagonals on a uniform mesh, a general
//     DAG with hanging nodes; off-rank (ghost) and periodic wrap-around inflow is lagged
//     (block Jacobi across ranks, which is what PETSc's bjacobi does too);
//   * one launch per level over the flat list of its (element, angle) pairs, one thread each
//     (k_pc_sweep_flat: full warps whatever the angles per quadrant; one block per element for
//     the generic nx != ny class): rhs = r - (inflow flux of the already computed z of the upwind columns,
//     through the same spatial / angular mortar tables as the apply), then the local solve
//     of the (column, angle) diagonal block  A_loc z = rhs,  nx ny unknowns;
//   * A_loc / wt = diag(mk) + cos(theta) (dy/2) Bx (x) Wy + sin(theta) (dx/2) Wx (x) By  with
//     the 1-D upwind operators B = -S + (outflow end node) -- with the extinction replaced by
//     its mid-range over the column it is a sum of Kronecker products of 1-D operators, which
//     the eigen-decomposition  W^-1 B = V diag(lambda) V^-1  (host, numpy; rtb200_pc_set_fd)
//     diagonalises:  12 n^3 FMAs per element instead of a dense LU; n_corr defect-correction
//     steps with the true nodal extinction follow (contraction factor
//     (kmax - kmin) / (kmax + kmin) < 1 by the mid-range choice);
//   * inflow-boundary nodes (get_masks.py:80-110) are removed from the local system: the 1-D
//     tables exist with and without the inflow end node.
//
// With this preconditioner GMRES converges in a few tens of iterations on meshes where the
// node-wise (M - S)^{-1} and the unpreconditioned iteration stagnate (free streaming between
// the scatterers needs one iteration per column otherwise).
#include "rtb200_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

static __constant__ double c_w[(RT_MAXN + 1) * RT_MAXN];
static __constant__ double c_s[(RT_MAXN + 1) * RT_MAXN * RT_MAXN];

struct SweepTab {                  // device pointers
    const int32_t* sig_qth0;       // [n_sig * 5] ranges in qth_pool: angles of quadrant q of a signature
    const int32_t* qth_pool;       // angular indices (position in the slab row)
    const int64_t* fd_off;         // [(RT_MAXN + 1) * 4] offsets (doubles) into fd_pool or -1
    const double* fd_pool;
    const double* col_kbar;        // [n_col] (dx dy / 4) * mid-range extinction of the column
};

// Everything a thread needs to know about its (column, quadrant) element, gathered on the host
// when the schedule is built: ONE load instead of a chain of dependent table look-ups
// (item -> column -> signature -> ...).
struct SweepElem {
    int64_t off;        // col_slab_off
    int64_t coef_off;   // col_coef_off
    int64_t tp0;        // sig_th0 of the column's signature
    int32_t c, q;
    int32_t nt, ntp;
    int32_t q0, nq;     // range of the quadrant's angles in qth_pool
    int32_t bdry, sg;
    int32_t cell0, nx;
    int32_t ny, pad;
    double hdx, hdy;    // dx / 2, dy / 2
};

struct SweepState {
    bool have_fd = false, have_sched = false, have_coef = false;
    SweepTab tab{};
    std::vector<void*> allocs;
    std::vector<int32_t> lvl_start;   // [n_levels * SW_NCLS + 1] ranges in items_dev: (level, kernel class)
    int32_t* items_dev = nullptr;     // (column << 2 | quadrant), level after level
    SweepElem* elems_dev = nullptr;   // the same sequence as records
    uint32_t* pairs_dev = nullptr;    // (element index << 8 | angle index in the quadrant): the flat work list
    std::vector<int64_t> pair_start;  // [n_levels * SW_NCLS + 1] ranges in pairs_dev
    int64_t n_items = 0;
    int block = 64;
    double rho_max = 0.0;             // max over columns of (kmax - kmin) / (kmax + kmin)
    double* kbar_dev = nullptr;
    double* rho_dev = nullptr;
};

namespace {

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cfma(const double2 a, const double2 b, const double2 c) {
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// 1-D tables of (n, variant): [lambda: n][V: n x n, row i][L: n x n, row alpha], complex
struct Fd1 {
    const double2* lam;
    const double2* V;
    const double2* L;
};
__device__ __forceinline__ Fd1 fd_tab(const SweepTab& st, const int n, const int v) {
    const double2* p = reinterpret_cast<const double2*>(st.fd_pool + st.fd_off[n * 4 + v]);
    return Fd1{p, p + n, p + n + n * n};
}

// z (+)= A~^{-1} b for one (column, angle): tensor-product fast diagonalisation.
//   kbwt = (dx dy / 4) kappa_mid wt,  ax = (dy/2) wt cos(theta),  ay = (dx/2) wt sin(theta)
template <int NF, bool ACC>
__device__ __forceinline__ void fd_solve(const int nx, const int ny, const Fd1 fx, const Fd1 fy,
                                         const double kbwt, const double ax, const double ay,
                                         const double (&b)[NF][NF], double (&z)[NF][NF]) {
    if (!ACC) {
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j) z[i][j] = 0.0;
    }
#pragma unroll 1
    for (int al = 0; al < nx; ++al) {
        double2 t1[NF], t2[NF];
        // t1[j] = sum_i Lx[al][i] b[i][j]
#pragma unroll
        for (int j = 0; j < NF; ++j) t1[j] = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            if (i < nx) {
                const double2 l = __ldg(&fx.L[al * nx + i]);
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    if (j < ny) {
                        t1[j].x = fma(l.x, b[i][j], t1[j].x);
                        t1[j].y = fma(l.y, b[i][j], t1[j].y);
                    }
            }
        }
        // t2[be] = (sum_j Ly[be][j] t1[j]) / (kbwt + ax lamx[al] + ay lamy[be])
        const double2 lx = __ldg(&fx.lam[al]);
#pragma unroll
        for (int be = 0; be < NF; ++be) {
            if (be < ny) {
                double2 s = make_double2(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    if (j < ny) s = cfma(__ldg(&fy.L[be * ny + j]), t1[j], s);
                const double2 ly = __ldg(&fy.lam[be]);
                const double dr = fma(ax, lx.x, fma(ay, ly.x, kbwt));
                const double di = fma(ax, lx.y, ay * ly.y);
                const double inv = 1.0 / fma(dr, dr, di * di);
                t2[be] = make_double2((s.x * dr + s.y * di) * inv, (s.y * dr - s.x * di) * inv);
            }
        }
        // t1[j] = sum_be Vy[j][be] t2[be]
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            if (j < ny) {
                double2 s = make_double2(0.0, 0.0);
#pragma unroll
                for (int be = 0; be < NF; ++be)
                    if (be < ny) s = cfma(__ldg(&fy.V[j * ny + be]), t2[be], s);
                t1[j] = s;
            }
        }
        // z[i][j] += Re(Vx[i][al] t1[j])
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            if (i < nx) {
                const double2 v = __ldg(&fx.V[i * nx + al]);
#pragma unroll
                for (int j = 0; j < NF; ++j)
                    if (j < ny) z[i][j] = fma(v.x, t1[j].x, fma(-v.y, t1[j].y, z[i][j]));
            }
        }
    }
}

// Column-level context of one (column, quadrant) element.
struct SweepCol {
    int c, nx, ny, sg, ntp, bdry, cell0;
    int64_t tp0, off;
    double hdx, hdy, kbar;
    const double* mk;
};

// One (column, quadrant, angle): right-hand side minus the inflow flux of the upwind columns,
// local solve, store.  N > 0: nx == ny == N at compile time; N == 0: any nx, ny <= RT_MAXN.
template <int N, bool CORR = true>
__device__ __forceinline__ void sweep_angle(const DevMesh& m, const SweepTab& st, const SweepCol& k,
                                            const int th, const bool couple, const int ncorr_in,
                                            const double* __restrict__ r, double* z) {
    constexpr int NF = N > 0 ? N : RT_MAXN;
    const int c = k.c;
    const int nx = N > 0 ? N : k.nx;
    const int ny = N > 0 ? N : k.ny;
    const int sg = k.sg;
    const int ntp = k.ntp;
    const int64_t tp0 = k.tp0;
    const int64_t off = k.off;
    const int bdry = k.bdry;
    const double hdx = k.hdx, hdy = k.hdy;
    const double* wx = c_w + nx * RT_MAXN;
    const double* wy = c_w + ny * RT_MAXN;
    const double* Sx = c_s + nx * RT_MAXN * RT_MAXN;
    const double* Sy = c_s + ny * RT_MAXN * RT_MAXN;
    const double* __restrict__ mk = k.mk;
    const double kbar = k.kbar;
    const int ncorr = CORR ? ncorr_in : 0;      // CORR = false: the defect-correction loop is compiled out
    {
        const int bits = m.th_bits[tp0 + th];
        const double wt = m.th_wt[tp0 + th], ct = m.th_ct[tp0 + th], sn = m.th_st[tp0 + th];
        const bool outx_hi = bits & 1, outy_hi = bits & 2;      // outflow through face 0 / face 1
        const int Fx = outx_hi ? 2 : 0, Fy = outy_hi ? 3 : 1;   // the two inflow faces
        const bool mx = (bdry >> Fx) & 1, my = (bdry >> Fy) & 1; // inflow through the domain boundary
        const int imx = outx_hi ? 0 : nx - 1, jmy = outy_hi ? 0 : ny - 1;   // inflow end nodes

        double b[NF][NF], zz[NF][NF];
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
                if (i < nx && j < ny) b[i][j] = r[off + (int64_t)(i * ny + j) * ntp + th];

        // ---- inflow flux of the upwind columns' z (boundary_convection_matrix.py:348-430) ---- //
        if (couple) {
            const int cell = m.th_cell[tp0 + th];
            const int rr = (bits >> 8) & 0xff;
            const int nth0 = m.cell_nth[k.cell0 + cell];
#pragma unroll
            for (int ff = 0; ff < 2; ++ff) {
                const int F = ff == 0 ? Fx : Fy;
                if ((bdry >> F) & 1) continue;
                const int n0 = (F & 1) ? nx : ny;
                const double half = (F & 1) ? hdx : hdy;
                const double* wq = (F & 1) ? wx : wy;
                const double wF = m.th_wF[F * m.n_theta_pool + tp0 + th];
                double fa[NF];
#pragma unroll
                for (int p = 0; p < NF; ++p) fa[p] = 0.0;
                for (int slot = 0; slot < 2; ++slot) {
                    const int fidx = c * 8 + F * 2 + slot;
                    if (m.face_nbr[fidx] < 0) continue;          // none / other rank / wrap: lagged
                    const double* src = z + m.face_base[fidx];
                    const int rs = m.face_rstride[fidx];
                    const int eo = m.face_exy[fidx];
                    const int ao = m.face_acpl[fidx];
                    if (eo < 0 && ao < 0) {
                        const double f = half * wF;
#pragma unroll
                        for (int p = 0; p < NF; ++p)
                            if (p < n0) fa[p] = f * wq[p] * src[(int64_t)p * rs + th];
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
                                const double* __restrict__ eth = m.eth_pool + acp[k * 3] + rr;
                                const double* tr = src + (int64_t)j1 * rs + acp[k * 3 + 1];
                                for (int a = 0; a < n1t; ++a) tmp = fma(eth[a * nth0], tr[a], tmp);
                            }
                        }
                        tmp *= half;
                        if (eo < 0) {
#pragma unroll
                            for (int p = 0; p < NF; ++p)
                                if (p == j1) fa[p] = fma(wq[p], tmp, fa[p]);
                        } else {
                            const double* __restrict__ E = m.exy_pool + eo;
#pragma unroll
                            for (int p = 0; p < NF; ++p)
                                if (p < n0) fa[p] = fma(E[j1 * n0 + p], tmp, fa[p]);
                        }
                    }
                }
                if (F & 1) {
#pragma unroll
                    for (int i = 0; i < NF; ++i)
#pragma unroll
                        for (int j = 0; j < NF; ++j)
                            if (i < nx && j == jmy) b[i][j] -= fa[i];
                } else {
#pragma unroll
                    for (int i = 0; i < NF; ++i)
#pragma unroll
                        for (int j = 0; j < NF; ++j)
                            if (i == imx && j < ny) b[i][j] -= fa[j];
                }
            }
        }
        // inflow-boundary nodes are not unknowns (get_masks.py:80-110)
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
                if ((mx && i == imx) || (my && j == jmy)) b[i][j] = 0.0;

        // ---- local solve ------------------------------------------------------------ //
        const Fd1 fx = fd_tab(st, nx, (outx_hi ? 0 : 1) + (mx ? 2 : 0));
        const Fd1 fy = fd_tab(st, ny, (outy_hi ? 0 : 1) + (my ? 2 : 0));
        const double ax = hdy * ct, ay = hdx * sn, kbwt = kbar * wt;
        fd_solve<NF, false>(nx, ny, fx, fy, kbwt, ax, ay, b, zz);
        for (int it = 0; it < ncorr; ++it) {
            // res = b - A_loc zz with the true nodal extinction
            double res[NF][NF];
            const double ox = outx_hi ? ax : -ax, oy = outy_hi ? ay : -ay;   // |.|: outflow weights
#pragma unroll
            for (int i = 0; i < NF; ++i)
#pragma unroll
                for (int j = 0; j < NF; ++j) {
                    if (i < nx && j < ny) {
                        double sx = 0.0, sy = 0.0;
#pragma unroll
                        for (int k = 0; k < NF; ++k)
                            if (k < nx) sx = fma(Sx[i * RT_MAXN + k], zz[k][j], sx);
#pragma unroll
                        for (int k = 0; k < NF; ++k)
                            if (k < ny) sy = fma(Sy[j * RT_MAXN + k], zz[i][k], sy);
                        double v = mk[i * ny + j] * wt * zz[i][j] - (wy[j] * ax * sx + wx[i] * ay * sy);
                        if (i == (outx_hi ? nx - 1 : 0)) v = fma(wy[j] * ox, zz[i][j], v);
                        if (j == (outy_hi ? ny - 1 : 0)) v = fma(wx[i] * oy, zz[i][j], v);
                        res[i][j] = ((mx && i == imx) || (my && j == jmy)) ? 0.0 : b[i][j] - v;
                    }
                }
            fd_solve<NF, true>(nx, ny, fx, fy, kbwt, ax, ay, res, zz);
        }
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
                if (i < nx && j < ny) z[off + (int64_t)(i * ny + j) * ntp + th] = zz[i][j];
    }
}

// One element: column c, quadrant q.  One thread per angle of the quadrant (generic class).
template <int N>
__device__ __forceinline__ void sweep_element(const DevMesh& m, const SweepTab& st, const int c,
                                              const int q, const bool couple, const int ncorr,
                                              const double* __restrict__ r, double* z) {
    const int sg = m.col_sig[c];
    const int nt = m.sig_nt[sg];
    SweepCol k;
    k.c = c; k.nx = m.col_nx[c]; k.ny = m.col_ny[c]; k.sg = sg; k.ntp = m.sig_ntp[sg];
    k.bdry = m.col_bdry[c]; k.cell0 = m.sig_cell0[sg];
    k.tp0 = m.sig_th0[sg]; k.off = m.col_slab_off[c];
    k.hdx = 0.5 * m.col_dx[c]; k.hdy = 0.5 * m.col_dy[c]; k.kbar = st.col_kbar[c];
    k.mk = m.coef + m.col_coef_off[c];
    const int q0 = st.sig_qth0[sg * 5 + q], nq = st.sig_qth0[sg * 5 + q + 1] - q0;
    if (q == 0 && k.ntp > nt)                               // pad entry of an odd row length
        for (int n = threadIdx.x; n < k.nx * k.ny; n += blockDim.x) z[k.off + (int64_t)n * k.ntp + nt] = 0.0;
    for (int t = threadIdx.x; t < nq; t += blockDim.x)
        sweep_angle<N>(m, st, k, st.qth_pool[q0 + t], couple, ncorr, r, z);
}

// Flat form: the (element, angle) pairs of a level are one work list, one thread each, so that
// every warp is full whatever the number of angles per quadrant (40 at the benchmark mesh: the
// one-block-per-element form left 24 of 64 lanes idle).  The element record replaces the
// column -> signature -> ... look-up chain.
template <int N, bool CORR>
__global__ void __launch_bounds__(128, CORR ? 2 : 3)
k_pc_sweep_flat(DevMesh m, SweepTab st, const SweepElem* __restrict__ elems, const uint32_t* __restrict__ pairs,
                const int n_pairs, int couple, int ncorr, const double* __restrict__ r, double* z) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_pairs) return;
    const uint32_t p = pairs[w];
    const SweepElem el = elems[p >> 8];
    const int t = (int)(p & 255u);
    SweepCol k;
    k.c = el.c; k.nx = el.nx; k.ny = el.ny; k.sg = el.sg; k.ntp = el.ntp; k.bdry = el.bdry; k.cell0 = el.cell0;
    k.tp0 = el.tp0; k.off = el.off; k.hdx = el.hdx; k.hdy = el.hdy; k.kbar = st.col_kbar[el.c];
    k.mk = m.coef + el.coef_off;
    if (t == 0 && el.q == 0 && el.ntp > el.nt)              // pad entry of an odd row length
        for (int n = 0; n < k.nx * k.ny; ++n) z[k.off + (int64_t)n * k.ntp + el.nt] = 0.0;
    sweep_angle<N, CORR>(m, st, k, st.qth_pool[el.q0 + t], couple, ncorr, r, z);
}

template <int N>
__global__ void __launch_bounds__(64)
k_pc_sweep(DevMesh m, SweepTab st, const int32_t* __restrict__ items, int couple, int ncorr,
           const double* __restrict__ r, double* z) {
    const int item = items[blockIdx.x];
    sweep_element<N>(m, st, item >> 2, item & 3, couple, ncorr, r, z);
}

constexpr int SW_NCLS = 6;                         // kernel classes: nx == ny == 2..6, else generic
__host__ __device__ inline int sweep_class(int nx, int ny) { return (nx == ny && nx >= 2 && nx <= 6) ? nx - 1 : 0; }

// per column: mid-range of the nodal extinction samples (times dx dy / 4) and the contraction
// factor of the defect correction
__global__ void k_col_kbar(DevMesh m, double* __restrict__ kbar, double* __restrict__ rho) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m.n_col) return;
    const int nx = m.col_nx[c], ny = m.col_ny[c];
    const double s = 0.25 * m.col_dx[c] * m.col_dy[c];
    const double* __restrict__ mk = m.coef + m.col_coef_off[c];
    double lo = 1e300, hi = 0.0;
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < ny; ++j) {
            const double k = mk[i * ny + j] / (s * c_w[nx * RT_MAXN + i] * c_w[ny * RT_MAXN + j]);
            lo = fmin(lo, k);
            hi = fmax(hi, k);
        }
    lo = fmax(lo, 0.0);
    hi = fmax(hi, 0.0);
    kbar[c] = s * 0.5 * (lo + hi);
    rho[c] = (hi + lo) > 0.0 ? (hi - lo) / (hi + lo) : 0.0;
}

// pairs[pair0[e] + t] = (e << 8 | t), t < nq of element e
__global__ void k_fill_pairs(const SweepElem* __restrict__ elems, const int64_t* __restrict__ pair0, const int ne,
                             uint32_t* __restrict__ pairs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int nq = elems[e].nq;
    uint32_t* dst = pairs + pair0[e];
    for (int t = 0; t < nq; ++t) dst[t] = ((uint32_t)e << 8) | (uint32_t)t;
}

template <typename T>
int dev_upload(rtb200_ctx* ctx, SweepState* S, const std::vector<T>& h, const T** d) {
    T* p = nullptr;
    RT_CUDA(ctx, cudaMalloc((void**)&p, sizeof(T) * std::max<size_t>(h.size(), 1)));
    S->allocs.push_back(p);
    if (!h.empty()) RT_CUDA(ctx, cudaMemcpy(p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    *d = p;
    return RTB200_OK;
}

int quadrant_of_bits(int bits) {
    switch (bits & 15) {
        case 3: return 0;     // outflow through faces 0, 1
        case 6: return 1;     // 1, 2
        case 12: return 2;    // 2, 3
        case 9: return 3;     // 3, 0
        default: return -1;
    }
}

// Topological levels of the (column, quadrant) elements under "reads the inflow trace of".
int build_schedule(rtb200_ctx* ctx, SweepState* S) {
    const int nc = ctx->dm.n_col;
    const int n_sig = (int)ctx->h_sig_nt.size();
    // angles per (signature, quadrant)
    std::vector<int32_t> qth0(n_sig * 5 + 1, 0), pool;
    int max_nq = 1;
    for (int s = 0; s < n_sig; ++s) {
        for (int q = 0; q < 4; ++q) {
            qth0[s * 5 + q] = (int32_t)pool.size();
            for (int th = 0; th < ctx->h_sig_nt[s]; ++th) {
                const int qq = quadrant_of_bits(ctx->h_th_bits[ctx->h_sig_th0[s] + th]);
                if (qq < 0)
                    return rt_fail(ctx, RTB200_EUNSUPPORTED,
                                   "transport preconditioner: an angular cell straddles two quadrants "
                                   "(angular level < 2)");
                if (qq == q) pool.push_back(th);
            }
            max_nq = std::max(max_nq, (int)pool.size() - qth0[s * 5 + q]);
        }
        qth0[s * 5 + 4] = (int32_t)pool.size();
    }
    // dependency graph
    const int64_t ni = (int64_t)nc * 4;
    std::vector<int32_t> indeg(ni, 0), level(ni, 0);
    std::vector<int64_t> adj0(ni + 1, 0);
    auto for_deps = [&](int c, int q, auto&& fn) {
        const int Fin[2] = {(q + 2) & 3, (q + 3) & 3};
        for (int f = 0; f < 2; ++f)
            for (int slot = 0; slot < 2; ++slot) {
                const int nb = ctx->h_face_nbr[c * 8 + Fin[f] * 2 + slot];
                if (nb >= 0) fn((int64_t)nb * 4 + q);
            }
    };
    for (int c = 0; c < nc; ++c)
        for (int q = 0; q < 4; ++q)
            for_deps(c, q, [&](int64_t d) { adj0[d + 1]++; indeg[(int64_t)c * 4 + q]++; });
    for (int64_t i = 0; i < ni; ++i) adj0[i + 1] += adj0[i];
    std::vector<int32_t> adj(adj0[ni]);
    {
        std::vector<int64_t> fill(adj0.begin(), adj0.end() - 1);
        for (int c = 0; c < nc; ++c)
            for (int q = 0; q < 4; ++q)
                for_deps(c, q, [&](int64_t d) { adj[fill[d]++] = c * 4 + q; });
    }
    std::vector<int32_t> order;
    order.reserve(ni);
    S->lvl_start.assign(1, 0);
    std::vector<int32_t> cur, nxt;
    for (int64_t i = 0; i < ni; ++i)
        if (indeg[i] == 0) cur.push_back((int32_t)i);
    auto cls_of = [&](int32_t i) { return sweep_class(ctx->h_col_nx[i >> 2], ctx->h_col_ny[i >> 2]); };
    while (!cur.empty()) {
        // inside a level: by kernel class, then memory order
        std::sort(cur.begin(), cur.end(), [&](int32_t a, int32_t b) {
            const int ca = cls_of(a), cb = cls_of(b);
            return ca != cb ? ca < cb : a < b;
        });
        size_t e = 0;
        for (int k = 0; k < SW_NCLS; ++k) {
            while (e < cur.size() && cls_of(cur[e]) == k) order.push_back(cur[e++]);
            S->lvl_start.push_back((int32_t)order.size());
        }
        nxt.clear();
        for (int32_t i : cur)
            for (int64_t e = adj0[i]; e < adj0[i + 1]; ++e)
                if (--indeg[adj[e]] == 0) nxt.push_back(adj[e]);
        cur.swap(nxt);
    }
    if ((int64_t)order.size() != ni)
        return rt_fail(ctx, RTB200_EUNSUPPORTED,
                       "transport preconditioner: the upwind dependency graph of the columns has a cycle");
    S->n_items = ni;
    S->block = 64;
    int rc;
    const int32_t* d32 = nullptr;
    if ((rc = dev_upload(ctx, S, order, &d32))) return rc;
    S->items_dev = const_cast<int32_t*>(d32);
    {
        // element records (see SweepElem) and the flat (element, angle) work list; the tables without
        // a host copy come back from the device (once per mesh)
        std::vector<int64_t> slab_off(std::max(nc, 1));
        std::vector<int32_t> sig_ntp(std::max(n_sig, 1)), sig_cell0(std::max(n_sig, 1));
        if (nc > 0)
            RT_CUDA(ctx, cudaMemcpy(slab_off.data(), ctx->dm.col_slab_off, sizeof(int64_t) * nc, cudaMemcpyDeviceToHost));
        if (n_sig > 0) {
            RT_CUDA(ctx, cudaMemcpy(sig_ntp.data(), ctx->dm.sig_ntp, sizeof(int32_t) * n_sig, cudaMemcpyDeviceToHost));
            RT_CUDA(ctx, cudaMemcpy(sig_cell0.data(), ctx->dm.sig_cell0, sizeof(int32_t) * n_sig, cudaMemcpyDeviceToHost));
        }
        std::vector<SweepElem> elems(order.size());
        std::vector<int64_t> pair0(order.size() + 1, 0);       // first pair of every element
        bool flat_ok = order.size() < (1u << 24);
        S->pair_start.assign(1, 0);
        size_t l = 0;
        for (size_t e = 0; e < order.size(); ++e) {
            while (l + 1 < S->lvl_start.size() && (int32_t)e >= S->lvl_start[l + 1]) {
                S->pair_start.push_back(pair0[e]);
                ++l;
            }
            const int c = order[e] >> 2, q = order[e] & 3, sg = ctx->h_col_sig[c];
            SweepElem& E = elems[e];
            E.off = slab_off[c];
            E.coef_off = ctx->h_col_coef_off[c];
            E.tp0 = ctx->h_sig_th0[sg];
            E.c = c; E.q = q;
            E.nt = ctx->h_sig_nt[sg]; E.ntp = sig_ntp[sg];
            E.q0 = qth0[sg * 5 + q]; E.nq = qth0[sg * 5 + q + 1] - E.q0;
            E.bdry = ctx->h_col_bdry[c]; E.sg = sg;
            E.cell0 = sig_cell0[sg]; E.nx = ctx->h_col_nx[c];
            E.ny = ctx->h_col_ny[c]; E.pad = 0;
            E.hdx = 0.5 * ctx->h_col_dx[c]; E.hdy = 0.5 * ctx->h_col_dy[c];
            if (E.nq > 255) flat_ok = false;
            pair0[e + 1] = pair0[e] + E.nq;
        }
        const int64_t n_pairs = pair0[order.size()];
        while (S->pair_start.size() < S->lvl_start.size()) S->pair_start.push_back(n_pairs);
        if (flat_ok && !order.empty()) {
            const SweepElem* de = nullptr;
            const int64_t* dp0 = nullptr;
            if ((rc = dev_upload(ctx, S, elems, &de))) return rc;
            if ((rc = dev_upload(ctx, S, pair0, &dp0))) return rc;
            S->elems_dev = const_cast<SweepElem*>(de);
            RT_CUDA(ctx, cudaMalloc((void**)&S->pairs_dev, sizeof(uint32_t) * std::max<int64_t>(n_pairs, 1)));
            S->allocs.push_back(S->pairs_dev);
            const int ne = (int)order.size();
            k_fill_pairs<<<(ne + 255) / 256, 256>>>(de, dp0, ne, S->pairs_dev);   // expanded on the device
            ctx->launches++;
            RT_CUDA(ctx, cudaGetLastError());
            RT_CUDA(ctx, cudaDeviceSynchronize());
        }
    }
    if ((rc = dev_upload(ctx, S, qth0, &S->tab.sig_qth0))) return rc;
    if ((rc = dev_upload(ctx, S, pool, &S->tab.qth_pool))) return rc;
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_w, ctx->h_lgl_w.data(), sizeof(double) * ctx->h_lgl_w.size()));
    RT_CUDA(ctx, cudaMemcpyToSymbol(c_s, ctx->h_lgl_s.data(), sizeof(double) * ctx->h_lgl_s.size()));
    RT_CUDA(ctx, cudaMalloc((void**)&S->kbar_dev, sizeof(double) * std::max(nc, 1)));
    S->allocs.push_back(S->kbar_dev);
    RT_CUDA(ctx, cudaMalloc((void**)&S->rho_dev, sizeof(double) * std::max(nc, 1)));
    S->allocs.push_back(S->rho_dev);
    S->tab.col_kbar = S->kbar_dev;
    S->have_sched = true;
    (void)max_nq;
    return RTB200_OK;
}

int prepare(rtb200_ctx* ctx, cudaStream_t st) {
    SweepState* S = ctx->sweep;
    if (!S || !S->have_fd)
        return rt_fail(ctx, RTB200_ESTATE,
                       "transport preconditioner: no 1-D tables (call rtb200_pc_set_fd after rtb200_set_mesh)");
    if (!S->have_sched) {
        int rc = build_schedule(ctx, S);
        if (rc) return rc;
    }
    if (!S->have_coef) {
        const int nc = ctx->dm.n_col;
        if (nc > 0) {
            k_col_kbar<<<(nc + 127) / 128, 128, 0, st>>>(ctx->dm, S->kbar_dev, S->rho_dev);
            ctx->launches++;
            RT_CUDA(ctx, cudaGetLastError());
            std::vector<double> rho(nc);
            RT_CUDA(ctx, cudaMemcpyAsync(rho.data(), S->rho_dev, sizeof(double) * nc, cudaMemcpyDeviceToHost, st));
            RT_CUDA(ctx, cudaStreamSynchronize(st));
            S->rho_max = *std::max_element(rho.begin(), rho.end());
        }
        S->have_coef = true;
    }
    return RTB200_OK;
}

}  // namespace

int rt_sweep_apply(rtb200_ctx* ctx, int pc_type, const double* r, double* z, cudaStream_t st) {
    int rc = prepare(ctx, st);
    if (rc) return rc;
    SweepState* S = ctx->sweep;
    if (S->n_items == 0) return RTB200_OK;
    // defect-correction steps of the local solves: contraction factor rho per step
    int ncorr = ctx->pc_ncorr;
    if (ncorr < 0) ncorr = S->rho_max <= 0.5 ? 0 : (S->rho_max <= 0.85 ? 1 : 2);
    const char* env = getenv("RTB200_PC_NCORR");
    if (env && env[0] >= '0' && env[0] <= '9') ncorr = env[0] - '0';
    // one launch per (level, kernel class); the element-wise block-Jacobi has no coupling, so all
    // levels of a class could run at once -- they are launched back to back on the same stream
    const int couple = pc_type == RTB200_PC_SWEEP ? 1 : 0;
    // flat (element, angle) work list unless RTB200_PC_SWEEP_FLAT=0 (or an element index / angle
    // count that does not fit the packed pair)
    const char* fe = getenv("RTB200_PC_SWEEP_FLAT");
    const bool flat = S->pairs_dev && !(fe && fe[0] == '0');
    for (size_t l = 0; l + 1 < S->lvl_start.size(); ++l) {
        const int i0 = S->lvl_start[l], cnt = S->lvl_start[l + 1] - i0;
        if (cnt == 0) continue;
        const int32_t* it = S->items_dev + i0;
#define SW_LAUNCH(N_) k_pc_sweep<N_><<<cnt, S->block, 0, st>>>(ctx->dm, S->tab, it, couple, ncorr, r, z)
#define SW_LAUNCH_FLAT(N_)                                                                                     \
    do {                                                                                                       \
        if (ncorr > 0)                                                                                         \
            k_pc_sweep_flat<N_, true><<<(np + 127) / 128, 128, 0, st>>>(ctx->dm, S->tab, S->elems_dev,         \
                                                                       S->pairs_dev + p0, np, couple, ncorr, r, z); \
        else                                                                                                   \
            k_pc_sweep_flat<N_, false><<<(np + 127) / 128, 128, 0, st>>>(ctx->dm, S->tab, S->elems_dev,        \
                                                                        S->pairs_dev + p0, np, couple, 0, r, z);  \
    } while (0)
        if (flat && l % SW_NCLS != 0) {
            const int64_t p0 = S->pair_start[l];
            const int np = (int)(S->pair_start[l + 1] - p0);
            if (np > 0) {
                switch ((int)(l % SW_NCLS)) {
                    case 1: SW_LAUNCH_FLAT(2); break;
                    case 2: SW_LAUNCH_FLAT(3); break;
                    case 3: SW_LAUNCH_FLAT(4); break;
                    case 4: SW_LAUNCH_FLAT(5); break;
                    default: SW_LAUNCH_FLAT(6); break;
                }
                ctx->launches++;
            }
            continue;
        }
        switch ((int)(l % SW_NCLS)) {
            case 1: SW_LAUNCH(2); break;
            case 2: SW_LAUNCH(3); break;
            case 3: SW_LAUNCH(4); break;
            case 4: SW_LAUNCH(5); break;
            case 5: SW_LAUNCH(6); break;
            default: SW_LAUNCH(0); break;
        }
#undef SW_LAUNCH
#undef SW_LAUNCH_FLAT
        ctx->launches++;
    }
    RT_CUDA(ctx, cudaGetLastError());
    return RTB200_OK;
}

void rt_sweep_destroy(rtb200_ctx* ctx) {
    if (!ctx->sweep) return;
    for (void* p : ctx->sweep->allocs) cudaFree(p);
    delete ctx->sweep;
    ctx->sweep = nullptr;
}

void rt_sweep_invalidate(rtb200_ctx* ctx) {
    if (ctx->sweep) ctx->sweep->have_coef = false;
}

extern "C" int rtb200_pc_set_fd(rtb200_ctx* ctx, const int64_t* fd_off, int32_t n_off,
                                const double* fd_pool, int64_t n_pool) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set (call rtb200_set_mesh first)");
    if (!fd_off || !fd_pool || n_off != (RT_MAXN + 1) * 4 || n_pool <= 0)
        return rt_fail(ctx, RTB200_EINVAL, "pc_set_fd: expected (RTB200_MAX_NXY + 1) * 4 table offsets");
    for (int i = 0; i < n_off; ++i) {
        const int n = i / 4;
        if (fd_off[i] >= 0 && ((fd_off[i] & 1) || fd_off[i] + 2 * (int64_t)(n + 2 * n * n) > n_pool))
            return rt_fail(ctx, RTB200_EINVAL, "pc_set_fd: table outside the pool");
    }
    // every order present in the mesh needs its four tables
    for (size_t c = 0; c < ctx->h_col_nx.size(); ++c)
        for (int v = 0; v < 4; ++v)
            if (fd_off[ctx->h_col_nx[c] * 4 + v] < 0 || fd_off[ctx->h_col_ny[c] * 4 + v] < 0)
                return rt_fail(ctx, RTB200_EINVAL, "pc_set_fd: missing table for an order of the mesh");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    rt_sweep_destroy(ctx);
    SweepState* S = new (std::nothrow) SweepState();
    if (!S) return rt_fail(ctx, RTB200_EINVAL, "out of host memory");
    ctx->sweep = S;
    std::vector<int64_t> off(fd_off, fd_off + n_off);
    std::vector<double> pool(fd_pool, fd_pool + n_pool);
    int rc;
    if ((rc = dev_upload(ctx, S, off, &S->tab.fd_off))) return rc;
    if ((rc = dev_upload(ctx, S, pool, &S->tab.fd_pool))) return rc;
    S->have_fd = true;
    return RTB200_OK;
}

extern "C" int rtb200_pc_set_type(rtb200_ctx* ctx, int pc_type, int n_corr) {
    if (!ctx) return RTB200_EINVAL;
    if (pc_type < RTB200_PC_NODE || pc_type > RTB200_PC_ELEMENT)
        return rt_fail(ctx, RTB200_EINVAL, "pc_set_type: unknown preconditioner");
    if (n_corr > 9) return rt_fail(ctx, RTB200_EINVAL, "pc_set_type: n_corr must be <= 9 (< 0: automatic)");
    ctx->pc_type = pc_type;
    ctx->pc_ncorr = n_corr;
    return RTB200_OK;
}

extern "C" int rtb200_pc_sweep_info(rtb200_ctx* ctx, int32_t* n_levels, int64_t* n_items) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "no mesh set");
    if (!ctx->sweep || !ctx->sweep->have_fd)
        return rt_fail(ctx, RTB200_ESTATE, "transport preconditioner: no 1-D tables (rtb200_pc_set_fd)");
    if (!ctx->sweep->have_sched) {
        int rc = build_schedule(ctx, ctx->sweep);
        if (rc) return rc;
    }
    if (n_levels) *n_levels = (int32_t)((ctx->sweep->lvl_start.size() - 1) / SW_NCLS);
    if (n_items) *n_items = ctx->sweep->n_items;
    return RTB200_OK;
}
