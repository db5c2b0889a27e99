/*
 * rtb200.h -- C ABI of librtb200.so, the B200 (sm_100a) transport-operator
 * library for hp-adaptive DGSEM radiative transfer (2-D space x 1-D angle).
 *
 * The reference (jasonltorchinsky/radiative-transfer-amr) is pure Python and
 * has no FFI layer: its "operator apply" is PETSc MatMult on the AIJ matrix
 * A = (BC - IC) + (M - S) assembled by src/rt (solve.py:49-54,68,100).  The
 * entry points below are what a binding for that path attaches to; each one
 * names the reference interface it replaces.  INTEGRATION.md shows the ctypes
 * stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no torch / C++ types in signatures;
 *   - every function returns 0 on success and a negative RTB200_E* code on
 *     failure, never throws or aborts; rtb200_last_error() gives the message;
 *   - the library never allocates caller-visible memory: vectors and the GMRES
 *     workspace are device buffers owned by the caller (PyTorch in this repo);
 *     the context owns only its private copies of the mesh tables;
 *   - one context per process per GPU, not thread-safe; all work is enqueued on
 *     the cudaStream_t passed as `void* stream`;
 *   - solver non-convergence is not an error: it is reported in
 *     rtb200_gmres_result.reason with PETSc's KSPConvergedReason codes.
 *
 * Vector layouts
 *   "dof order"  : the reference's global vector (dg/projection/to_vector.py:
 *                  16-42): columns in sorted-key order, cells in sorted-key
 *                  order, then (i, j, a) with the angular index fastest.
 *   "slab order" : the library's internal layout: per column one slab
 *                  [node = i*ny + j][theta], theta running over the column's
 *                  cells in sorted-key order, row stride padded to an even
 *                  count; pad entries are zero.  rtb200_to_slab / _from_slab
 *                  convert.  apply/gmres work on slab-order device vectors of
 *                  rtb200_flatmesh.n_slab doubles.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB200_ABI_VERSION 7

/* error codes */
#define RTB200_OK            0
#define RTB200_EINVAL       -1   /* bad argument / inconsistent tables     */
#define RTB200_ECUDA        -2   /* CUDA runtime error (see last_error)    */
#define RTB200_ESTATE       -3   /* call out of order (e.g. no mesh set)   */
#define RTB200_EUNSUPPORTED -4   /* size outside the compiled limits       */
#define RTB200_ENCCL        -5   /* NCCL error                             */

/* operator terms, with the sign they carry in A = (BC - IC) + (M - S)
 * (reference: src/rt/solve.py:52) */
#define RTB200_TERM_MASS   1u    /* + M   mass_matrix.py:105-180                    */
#define RTB200_TERM_SCAT   2u    /* - S   scattering_matrix.py / preconditioner_matrix.py:116-267 */
#define RTB200_TERM_ICONV  4u    /* - IC  interior_convection_matrix.py:110-250     */
#define RTB200_TERM_BCONV  8u    /* + BC  boundary_convection_matrix.py:160-540     */
#define RTB200_TERM_ALL   15u

/* limits of the compiled kernels */
#define RTB200_MAX_NXY     12    /* max nodes per spatial direction in a column */
#define RTB200_MAX_NTH     16    /* max angular nodes per cell                   */

typedef struct rtb200_ctx rtb200_ctx;

/*
 * Flattened hp-mesh layout, built on the host
 * once per AMR step from the reference's Mesh object
 * (dg/mesh/class_Mesh.py:12-90) by rt_b200.dg.flatmesh.  All pointers are HOST
 * pointers; rtb200_set_mesh copies them to the device.
 *
 * Column c (position in sorted-key order, restricted to this rank's block):
 *   nx, ny, angular signature, sizes, slab/dof offsets, boundary faces.
 * Angular signature s: the ordered list of (cell key, n_theta) of a column;
 *   columns with equal signatures share every angular table and G.
 * Face tables are indexed [c*8 + F*2 + slot], F: 0=+x 1=+y 2=-x 3=-y,
 *   slot 0/1 = first/second neighbour column across the face
 *   (Column.nhbr_keys[F], counter-clockwise order).
 */
typedef struct rtb200_flatmesh {
    int32_t abi_version;
    int32_t n_col;              /* columns owned by this rank                    */
    int32_t n_sig;              /* distinct angular signatures                   */
    int32_t n_sigcell;          /* total cells over all signatures               */
    int64_t n_dof;              /* doubles in a dof-order vector (owned columns) */
    int64_t n_slab;             /* doubles in a slab-order vector                */
    int64_t n_ghost;            /* doubles in the halo (ghost trace) buffer      */

    /* per column [n_col] */
    const int32_t* col_nx;
    const int32_t* col_ny;
    const int32_t* col_sig;
    const int32_t* col_bdry;    /* bit F set: face F lies on the domain boundary */
    const int64_t* col_dof_off; /* first dof-order index                         */
    const int64_t* col_slab_off;/* first slab-order index (even)                 */
    const int64_t* col_coef_off;/* first index into coef (2*nx*ny doubles/column)*/
    const double*  col_dx;
    const double*  col_dy;

    /* per (column, face, slot) [n_col*8] */
    const int64_t* face_base;   /* slab index of the neighbour's first face row
                                   (node j=0, theta 0); >= 0: in x;
                                   <= -2: -(v+2) indexes the ghost buffer;
                                   -1: no neighbour in this slot                 */
    const int32_t* face_rstride;/* doubles between consecutive face nodes there  */
    const int32_t* face_exy;    /* offset into exy_pool of E[n1][n0] (row j, col q) */
    const int32_t* face_n1;     /* face nodes of the neighbour column            */
    const int32_t* face_acpl;   /* offset into acpl_pool                         */

    /* per signature [n_sig] */
    const int32_t* sig_nt;      /* angular DoFs of a column                      */
    const int32_t* sig_ntp;     /* row stride (nt rounded up to even)            */
    const int32_t* sig_ncell;
    const int32_t* sig_cell0;   /* first entry in the sigcell arrays             */
    const int64_t* sig_th0;     /* first entry in the per-theta arrays           */
    const int64_t* sig_g_off;   /* first entry in g_pool: G[theta][theta'] row-major,
                                   zero-padded to [8*ceil(nt/8)][8*ceil(nt/8)]   */

    /* per signature cell [n_sigcell], sorted-key order inside a signature */
    const int32_t* cell_nth;
    const int32_t* cell_thoff;  /* first theta index of the cell in its column   */

    /* per theta, pooled over signatures (ntp entries each, pad entries zero) */
    int64_t n_theta_pool;
    const double*  th_wt;       /* (dtheta/2) * w_a                              */
    const double*  th_ct;       /* th_wt * cos(theta)                            */
    const double*  th_st;       /* th_wt * sin(theta)                            */
    const double*  th_wF;       /* [4][n_theta_pool]: th_wt * cos(theta - F pi/2)*/
    const int32_t* th_cell;     /* cell index inside the signature               */
    const int32_t* th_bits;     /* bits 0-3: face F is OUTFLOW for the cell's
                                   quadrant; bits 4-7: face F is INFLOW for the
                                   mask (get_masks.py:80-110); bits 8-15: index
                                   of the dof inside its cell                    */

    /* pools */
    int64_t n_g_pool;     const double*  g_pool;    /* per signature G, [theta][theta']   */
    int64_t n_exy_pool;   const double*  exy_pool;  /* spatial mortars E^x / E^y          */
    int64_t n_eth_pool;   const double*  eth_pool;  /* (dtheta0/2) * E^theta blocks [a][r]*/
    int64_t n_acpl_pool;  const int32_t* acpl_pool; /* per (sig0, sig1, F): [ncell0][2][3]
                                                       = (eth offset, theta' start, n_theta1) */

    /* per n = 0..RTB200_MAX_NXY: LGL weights and S[p][i] = phi_p'(xi_i) w_i */
    const double*  lgl_w;       /* [(MAX_NXY+1)][MAX_NXY]                        */
    const double*  lgl_s;       /* [(MAX_NXY+1)][MAX_NXY][MAX_NXY]               */

    /* coefficient samples: per column, mk[node] then ms[node]
       mk = (dx dy/4) w_i w_j kappa(x_i, y_j),  ms = (dx dy/4) w_i w_j sigma(x_i, y_j) */
    int64_t n_coef;       const double*  coef;

    /* halo exchange plan (multi-GPU; all zero / NULL on one GPU) */
    int32_t n_peer;             /* ranks this rank exchanges traces with         */
    const int32_t* peer_rank;   /* [n_peer]                                      */
    const int64_t* peer_send0;  /* [n_peer+1] ranges in send_src                 */
    const int64_t* peer_recv0;  /* [n_peer+1] ranges in the ghost buffer         */
    int64_t n_send;       const int64_t* send_src;  /* slab index of every double to send */

    /* block-circulant scattering tables (optional; n_circ_pool = 0: none).
       A signature whose cells all have the same width and order (NC cells of P
       nodes) and whose G is block circulant in the cell index,
       G[(K,r),(K',a)] = C[(K-K') mod NC][r][a] (any Phi(theta - theta')), may
       carry the block spectrum of C: the kernels then apply the scattering
       sum with FFTs over the cell index instead of the dense product.
       sig_circ_off[s] < 0: signature s has no such table.  Layout of one table:
       [NC/2][P][P] complex (re, im) doubles, slot 0 = (Chat[0], Chat[NC/2]) / NC,
       slot k = Chat[k] / (2 NC)  (rt/tables.py:circulant_spectrum). */
    const int32_t* sig_circ_nc;   /* [n_sig] NC (power of two, 4..64) or 0        */
    const int32_t* sig_circ_p;    /* [n_sig] P                                    */
    const int64_t* sig_circ_off;  /* [n_sig] even offset into circ_pool or -1     */
    int64_t n_circ_pool;  const double*  circ_pool;

    /* per (column, face, slot) [n_col*8]: LOCAL index of the neighbour column across the face
       (Column.nhbr_keys, class_Mesh.py:39-52) when it is owned by this rank and geometrically
       adjacent; -1: none; -2: owned by another rank (ghost trace); -3: periodic wrap-around.
       Used by the transport-sweep preconditioner to order the columns upwind. */
    const int32_t* face_nbr;
} rtb200_flatmesh;

typedef struct rtb200_gmres_params {
    int32_t restart;        /* Krylov dimension before restart (reference: 50, solve.py:90) */
    int32_t max_it;         /* reference: 5000 (solve.py:86)                                */
    double  rtol;           /* reference: consts.EPS = 1e-15                                */
    double  atol;           /* reference: consts.EPS = 1e-15                                */
    double  divtol;         /* reference: sqrt(consts.INF)                                  */
    int32_t orth;           /* 0: classical Gram-Schmidt, refine if needed (DGKS);
                               1: always two passes; 2: never refine (PETSc default)        */
    int32_t pc;             /* RTB200_PC_*: 0 none, 1 node-wise (M - S)^-1, 2 transport sweep,
                               3 element-wise block-Jacobi (the sweep without upwind coupling) */
    int32_t x0_nonzero;     /* reference: 1 (solve.py:97)                                   */
    int32_t record_history; /* 1: write residual norms to hist                              */
    int32_t pc_side;        /* 0: left (KSPGMRES default; preconditioned residual norms);
                               1: right, flexible (KSPFGMRES, solve.py:83: the z_k = M^-1 v_k are
                                  stored, unpreconditioned residual norms)                   */
} rtb200_gmres_params;

typedef struct rtb200_gmres_result {
    int32_t reason;         /* PETSc code: 2 RTOL, 3 ATOL, -3 ITS, -4 DTOL, -5 BREAKDOWN     */
    int32_t iterations;
    double  res_norm;       /* last (preconditioned) residual norm estimate                 */
    double  res0_norm;      /* initial residual norm                                        */
    double  b_norm;
    double  true_res_norm;  /* ||b - A x|| recomputed at exit                               */
    double  sigma_max;      /* extreme singular values of the Hessenberg matrix, the        */
    double  sigma_min;      /*   estimate behind KSPComputeExtremeSingularValues            */
    int64_t n_apply;        /* operator applications performed                              */
} rtb200_gmres_result;

/* ---- life cycle -------------------------------------------------------- */
int  rtb200_abi_version(void);
int  rtb200_create(int device, rtb200_ctx** out);
void rtb200_destroy(rtb200_ctx* ctx);
const char* rtb200_last_error(const rtb200_ctx* ctx);

/* ---- element tables (replaces the five calc_col_matrix assemblers:
 *      preconditioner_matrix.py:116, interior_convection_matrix.py:110,
 *      boundary_convection_matrix.py:160, mass/scattering_matrix.py) ---- */
int  rtb200_set_mesh(rtb200_ctx* ctx, const rtb200_flatmesh* mesh);
/* new kappa/sigma samples on an unchanged mesh (same layout as flatmesh.coef) */
int  rtb200_set_coeffs(rtb200_ctx* ctx, const double* coef_host, int64_t n_coef);

/* ---- layout conversion (dg/projection/to_vector.py:16-42,
 *      from_vector.py:11-40 define dof order) --------------------------- */
int  rtb200_to_slab(rtb200_ctx* ctx, const double* x_dof_dev, double* x_slab_dev, void* stream);
int  rtb200_from_slab(rtb200_ctx* ctx, const double* x_slab_dev, double* x_dof_dev, void* stream);

/* keep = 1: zero the inflow-boundary entries (x <- mask .* x);
 * keep = 2: zero the interior entries        (x <- (1 - mask) .* x)
 * (dg/matrix/get_masks.py:19-129, split_vector.py:15-48) */
int  rtb200_mask(rtb200_ctx* ctx, double* x_slab_dev, int keep, void* stream);

/* ---- operator apply: y = sum of the selected terms applied to x
 *      (replaces PETSc MatMult on sys_mat / A_ii / A_ib, solve.py:52-54,68,100).
 *      interior_only = 1: rows of inflow-boundary DoFs are written as zero
 *      (A[I, :] x); with x zero on the boundary this is A_ii x, with x zero
 *      in the interior it is A_ib g.  x and y must not alias. ------------ */
int  rtb200_apply(rtb200_ctx* ctx, unsigned terms, int interior_only,
                  const double* x_slab_dev, double* y_slab_dev, void* stream);

/* ---- Krylov solve of A_ii x = b (replaces KSPSolve, solve.py:80-100).
 *      b, x: slab order, zero on boundary entries and pads.
 *      work: rtb200_gmres_work_bytes() bytes of device memory.
 *      hist: host array of max_it+1 doubles or NULL. -------------------- */
int64_t rtb200_gmres_work_bytes(const rtb200_ctx* ctx, int restart, int pc_side);
int  rtb200_gmres(rtb200_ctx* ctx, const double* b_dev, double* x_dev,
                  const rtb200_gmres_params* prm, rtb200_gmres_result* res,
                  double* hist_host, void* work_dev, int64_t work_bytes, void* stream);

/* preconditioner kinds (rtb200_gmres_params.pc, rtb200_pc_set_type) */
#define RTB200_PC_NONE     0
#define RTB200_PC_NODE     1     /* node-wise block-Jacobi: (M - S)^-1 per (column, spatial node)      */
#define RTB200_PC_SWEEP    2     /* transport sweep: (M + BC - IC)^-1 by upwind-ordered block
                                    Gauss-Seidel over (column, angle) elements -- what PETSc's
                                    bjacobi + ILU(0) (solve.py:93-95) approximates per rank        */
#define RTB200_PC_ELEMENT  3     /* element-wise block-Jacobi: the (column, angle) diagonal blocks
                                    of M + BC - IC only (north star, subsystem 4)                  */

/* ---- preconditioner: z = mask . (M - S)^{-1} r, the inverse of the reference's
 *      "preconditioner matrix" (preconditioner_matrix.py:22-267, block diagonal
 *      over (column, spatial node)) through the block-circulant structure of the
 *      scattering kernel on uniform angular meshes (>= 8 cells), through the eigen
 *      tables of rtb200_pc_set_eig otherwise (RTB200_EUNSUPPORTED without them).  rtb200_gmres applies it from the left when
 *      params.pc = 1.  x and y must not alias. --------------------------- */
int  rtb200_pc_apply(rtb200_ctx* ctx, const double* x_slab_dev, double* y_slab_dev, void* stream);
/* eigen tables of the same preconditioner for angular signatures WITHOUT a block-circulant
 * table (hp-refined angular meshes, fewer than 8 cells): per signature
 * [lambda: nt][W: nt x nt][W^T: nt x nt] with D^{-1/2} G D^{-1/2} = Q diag(lambda) Q^T,
 * W = D^{-1/2} Q, D = diag(th_wt), so that (mk D - ms G)^{-1} = W diag(1/(mk - ms lambda)) W^T
 * (symmetric scattering kernels; built by rt_b200.rt.tables.pc_eig_tables -- the batched
 * factorisation of the node blocks reduced to one eigen-decomposition per signature).
 * sig_eig_off[n_sig]: offset into eig_pool or -1.  HOST pointers, copied.  Call after
 * rtb200_set_mesh. */
int  rtb200_pc_set_eig(rtb200_ctx* ctx, const int64_t* sig_eig_off, int32_t n_sig,
                       const double* eig_pool, int64_t n_pool);
/* extinction floor kappa_eff = kappa + shift / min(dx, dy) used inside the preconditioner
 * (vacuum regions make M - S singular); default 1, 0 = exact (M - S)^{-1} */
int  rtb200_pc_set_shift(rtb200_ctx* ctx, double shift);

/* ---- transport-aware preconditioners (RTB200_PC_SWEEP / RTB200_PC_ELEMENT).
 *      z = mask . T~^{-1} r with T~ = M~ + BC - IC: the streaming part of the operator
 *      (mass_matrix.py, interior_convection_matrix.py, boundary_convection_matrix.py) with the
 *      extinction of a column replaced by the mid-range of its nodal samples.  T~ is block lower
 *      triangular when the (column, quadrant) elements are ordered upwind (the inflow faces of a
 *      quadrant only see columns that come earlier), so T~^{-1} r is ONE sweep: the library
 *      orders the elements topologically (levels of independent elements, one launch per level)
 *      and solves each (column, angle) diagonal block -- nx ny unknowns -- by tensor-product
 *      fast diagonalisation of the 1-D upwind operators, followed by n_corr defect-correction
 *      steps with the true nodal extinction.  Off-rank (ghost) and periodic wrap-around inflow is
 *      lagged (block Jacobi across ranks, like PETSc's bjacobi).
 *      rtb200_pc_set_fd uploads the 1-D tables (built by rt_b200.rt.tables.fd_tables): for every
 *      n = 1..RTB200_MAX_NXY and variant v = (flow sign < 0) + 2 (inflow end node removed)
 *      at fd_off[n * 4 + v]:  [lambda: n][V: n x n, row i][L = V^-1 W^-1: n x n, row alpha],
 *      complex (re, im) doubles, for the eigen-decomposition W^-1 B = V diag(lambda) V^-1 of the
 *      1-D operator B = -S + (outflow end node).  HOST pointers, copied.
 *      rtb200_pc_set_type selects what rtb200_pc_apply computes (default RTB200_PC_NODE). --- */
int  rtb200_pc_set_fd(rtb200_ctx* ctx, const int64_t* fd_off, int32_t n_off,
                      const double* fd_pool, int64_t n_pool);
int  rtb200_pc_set_type(rtb200_ctx* ctx, int pc_type, int n_corr);
/* sweep statistics after the first use: levels (launches per sweep), elements */
int  rtb200_pc_sweep_info(rtb200_ctx* ctx, int32_t* n_levels, int64_t* n_items);

/* ---- multi-GPU: one process per GPU; columns split like the reference's
 *      np.array_split(sorted keys, size)[rank] (mass_matrix.py:50-51).
 *      The communicator is created inside the library from an ncclUniqueId
 *      that rank 0 obtains with rtb200_nccl_unique_id and the host broadcasts
 *      (torch.distributed in this repo). -------------------------------- */
int  rtb200_nccl_unique_id(char* id128);                 /* 128-byte buffer */
int  rtb200_comm_init(rtb200_ctx* ctx, const char* id128, int rank, int nranks);
int  rtb200_halo_exchange(rtb200_ctx* ctx, const double* x_slab_dev, void* stream);

/* ---- AMR error-indicator reductions on the solution the solver leaves on the
 *      device (replaces the Python loops of
 *      src/amr/Error_Indicator/error/error_angular_jump.py:14-96 and the two
 *      integrals behind the hp steering, src/dg/projection/cell_intg_xy.py:14-41
 *      and col_intg_th.py:14-43).  One launch, one read of the slab.
 *      Tables (DEVICE pointers, owned by the caller):
 *        cell_lo / cell_up [n_sigcell]  position inside its signature of the cell
 *                                       below / above in theta (periodic;
 *                                       Cell.nhbr_keys, class_Mesh.py:27)
 *        lag_bot / lag_top [(MAX_NTH+1)][MAX_NTH]  Lagrange basis on the n
 *                                       Gauss-Legendre nodes at -1 / +1
 *        lgl_w             [(MAX_NXY+1)][MAX_NXY]  LGL weights (as flatmesh.lgl_w)
 *        col_cell_off [n_col]  first entry of column c in cell_err2
 *        col_th_off   [n_col]  first entry of column c in cell_intg_xy
 *        max_cell_nodes        max over columns of (cells x spatial nodes)
 *        max_slab_doubles      max over columns of nx * ny * row stride (the slab is staged
 *                              in shared memory when it fits); 0: never stage
 *      Outputs (device):
 *        cell_err2    [sum of cells]    squared cell indicator, cells in sorted-key order
 *        cell_intg_xy [sum of n_theta]  (dx dy/4) sum_ij w_i w_j u, theta in slab order
 *        col_intg_th  [n_coef / 2]      sum_theta (dtheta/2) w_a u, node = i*ny + j,
 *                                       column c at col_coef_off[c] / 2 ------------- */
int  rtb200_err_angular_jump(rtb200_ctx* ctx, const double* u_slab_dev,
                             const int32_t* cell_lo_dev, const int32_t* cell_up_dev,
                             const double* lag_bot_dev, const double* lag_top_dev,
                             const double* lgl_w_dev, const int64_t* col_cell_off_dev,
                             const int64_t* col_th_off_dev, int32_t max_cell_nodes,
                             int32_t max_slab_doubles, double* cell_err2_dev,
                             double* cell_intg_xy_dev, double* col_intg_th_dev, void* stream);

/* ---- spatial-jump indicator (replaces src/amr/Error_Indicator/error/
 *      error_spatial_jump.py:30-252): per column, the squared difference of the two
 *      traces across every face, interpolated to a common quadrature (smaller
 *      interval, larger node count, in space and in angle) and integrated;
 *      col_err2[c] = sum / (perimeter * 2 pi).  One launch; single rank.
 *      Tables (DEVICE pointers, built by rt_b200.amr.spatial_jump):
 *        sp_off / ap_off / nb [n_col*8]  per (column, face, slot): offset of the spatial
 *                  block [S0: n0 x nB][S1: n1 x nB][wB: nB][dA] in pool (-1: no
 *                  neighbour), offset of the (signature, neighbour signature) table in
 *                  apair, nB (negative: S0 = S1 = identity, same nodes on both sides)
 *        apair     [ncell0][2][4] per signature pair: (angular block offset in pool | -1,
 *                  first theta of the neighbour cell, nth1, nH (negative: T0 = T1 =
 *                  identity)); angular block =
 *                  [T0: nth0 x nH][T1: nth1 x nH][wH: nH][dG]
 *      Output: col_err2 [n_col] (device) ------------------------------------ */
int  rtb200_err_spatial_jump(rtb200_ctx* ctx, const double* u_slab_dev,
                             const int32_t* sp_off_dev, const int32_t* ap_off_dev,
                             const int32_t* nb_dev, const double* pool_dev,
                             const int32_t* apair_dev, double* col_err2_dev, void* stream);

/* ---- high-resolution error (replaces the cell loops of src/amr/Error_Indicator/error/
 *      error_high_resolution.py:83-216): ctx_lo / ctx_hi hold a mesh and a p-refined copy of
 *      it (same columns and cells, more nodes: ref_mesh(form = "p")); the low-resolution
 *      solution is interpolated to the high-resolution nodes and compared there:
 *        cell_err2[cells] = (dx dy dth / 8) sum w (u_hi - I u_lo)^2      (:194-205)
 *        col_hr2[n_col]   = sum over the column's cells of (dx dy / 4)(dth / 2) sum w u_hi^2 (:83-113)
 *      Tables (DEVICE pointers): lgl_ip_off [(MAX_NXY+1)^2] / lg_ip_off [(MAX_NTH+1)^2] offsets
 *      into ip_pool of the [n][m] Lagrange interpolation from n to m nodes; lgl_w
 *      [(MAX_NXY+1)][MAX_NXY], lg_w [(MAX_NTH+1)][MAX_NTH] weights; cell_dth [n_sigcell of
 *      ctx_lo] cell widths; col_cell_off [n_col]; max_cell_dofs = largest low-resolution cell. */
int  rtb200_err_high_resolution(rtb200_ctx* ctx_lo, const double* u_lo_slab_dev,
                                rtb200_ctx* ctx_hi, const double* u_hi_slab_dev,
                                const int32_t* lgl_ip_off_dev, const int32_t* lg_ip_off_dev,
                                const double* ip_pool_dev, const double* lgl_w_dev,
                                const double* lg_w_dev, const double* cell_dth_dev,
                                const int64_t* col_cell_off_dev, int32_t max_cell_dofs,
                                double* cell_err2_dev, double* col_hr2_dev, void* stream);

/* ---- instrumentation -------------------------------------------------- */
int64_t rtb200_launch_count(const rtb200_ctx* ctx);      /* kernels launched so far */
/* Which kernel serves which columns of the mesh set with rtb200_set_mesh: one line per launch
 * bucket, "kernel<template arguments> columns=<n> groups=<n> smem=<bytes>", newline separated,
 * NUL terminated, truncated to buf_len.  Returns the number of buckets (< 0: error).  The
 * reference has one code path (PETSc MatMult on the assembled AIJ matrix, src/rt/solve.py:68);
 * here the path depends on the column class, and Problem.solve reports it in
 * matrix_info["apply_paths"].  Lines of the persistent circulant kernel end with
 * "halo_stops=<n>": blocks of multi-GPU applies that found the halo flag down at their first
 * ghost-reading column so far (overlapped pass only, RTB200_HALO_OVERLAP=1; 0 otherwise). */
int  rtb200_apply_path(const rtb200_ctx* ctx, char* buf, int32_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
