// Internal declarations shared by the translation units of librtb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "rtb200.h"

#define RT_MAXN RTB200_MAX_NXY

// Device-side view of rtb200_flatmesh (all pointers are device pointers).
struct DevMesh {
    int32_t n_col;
    int64_t n_dof, n_slab;
    const int32_t *col_nx, *col_ny, *col_sig, *col_bdry;
    const int64_t *col_dof_off, *col_slab_off, *col_coef_off;
    const double *col_dx, *col_dy;
    const int64_t* face_base;
    const int32_t *face_rstride, *face_exy, *face_n1, *face_acpl;
    const int32_t *sig_nt, *sig_ntp, *sig_ncell, *sig_cell0;
    const int64_t *sig_th0, *sig_g_off;
    const int32_t *cell_nth, *cell_thoff;
    int64_t n_theta_pool;
    const double *th_wt, *th_ct, *th_st, *th_wF;
    const int32_t *th_cell, *th_bits;
    const double *g_pool, *exy_pool, *eth_pool;
    const int32_t* acpl_pool;
    const double* coef;
    const int32_t* face_nbr;
};

// A launch bucket: columns with nx == ny == n (n = 0: anything else).
struct Bucket {
    int n = 0;                 // template selector
    int count = 0;
    int n_groups = 0;          // thread blocks: groups of same-class columns
    int max_ntp = 0;
    int ntps = 0;              // shared-memory row stride (doubles)
    size_t smem = 0;           // bytes of dynamic shared memory
    int32_t* cols_dev = nullptr;        // column indices, group after group
    int32_t* grp_first_dev = nullptr;   // [n_groups] first entry in cols_dev
    int32_t* grp_count_dev = nullptr;   // [n_groups] columns in the group
    // block-circulant scattering (uniform angular mesh): NC cells of P nodes; 0 = dense
    int circ_nc = 0, circ_p = 0, circ_sig = 0;
    int32_t* cols_int_dev = nullptr;    // multi-GPU: columns that read no ghost trace ...
    int32_t* cols_halo_dev = nullptr;   // ... and those that do (rt_apply_with_halo)
    int32_t* cols_all_dev = nullptr;    // both lists back to back (single launch with a device-side wait)
    int count_int = 0, count_halo = 0;
    bool circ_herm = false;             // every block of the spectrum is Hermitian (symmetric Phi)
    bool circ_conforming = false;       // no mortar face in any column of the bucket
    bool hp = false;                    // dense columns served by k_apply_hp (apply_hp.cu)
    const void* hp_groups_dev = nullptr; // [n_groups] HpGroup records (apply_hp.h)
    // multi-GPU: the first n_free groups (hp) / columns (circulant) read no ghost trace and are
    // launched while the halo exchange is in flight (-1: the bucket is not split)
    int n_free = -1;
    const double* circ_tab_dev = nullptr;
    // eigen-table preconditioner (precond_eig.cuh): shared-memory sizing, signatures present
    size_t pc_max_rows_nt = 0;          // max over columns of nn * nt
    int pc_max_nn = 0;
    std::vector<int> sigs;
    std::vector<double> circ_tab_host;  // same table, host copy (kernel parameter of k_apply_circ7)
};

struct NcclApi;                // resolved with dlopen in comm.cu

struct rtb200_ctx {
    int device = 0;
    std::string err;
    bool have_mesh = false;
    DevMesh dm{};
    std::vector<void*> allocs; // private device allocations (tables)
    std::vector<Bucket> buckets;
    std::vector<int> bucket_order;              // launch order (largest first)
    std::vector<cudaStream_t> side_streams;     // buckets of an hp mesh run concurrently
    std::vector<cudaEvent_t> ev_join;
    cudaEvent_t ev_fork = nullptr;
    cudaStream_t comm_stream = nullptr;         // halo exchange overlapped with interior columns
    cudaEvent_t ev_halo = nullptr;
    bool overlap_ready = false;
    bool overlap_multi = false;        // several buckets: ghost-free parts overlap the exchange
    bool halo_fused = true;            // RTB200_HALO_FUSED=0: two launches (interior, then ghost-reading columns)
    unsigned* halo_flag_dev = nullptr; // raised to the apply's epoch by the communication stream
    unsigned halo_epoch = 0;
    int* halo_resume_dev = nullptr;    // per block: first column the first launch of a split apply left undone
    double* coef_dev = nullptr;
    int64_t n_coef = 0;
    const double* circ_pool_dev = nullptr;
    bool use_circ = true;          // RTB200_DENSE_SCATTERING=1 in the environment disables it
    int circ_version = 7;          // RTB200_CIRC_VERSION=6 selects the one-block-per-column kernel
    // k_apply_circ7 column-schedule counters, one pair per (bucket column list, stream): launches on one
    // stream are ordered, launches of the same bucket on different streams (the two e2e pipelines) must
    // not share a counter
    std::map<std::pair<const void*, void*>, unsigned*> sched_slots;
    cudaEvent_t ev_packed = nullptr;            // comm stream: traces packed, NCCL next (rt_apply_with_halo)
    int halo_sm_reserve = 0;       // RTB200_HALO_SM_RESERVE: SMs the persistent apply leaves to the halo exchange's kernels
    bool use_hp = true;            // RTB200_HP_KERNEL=old keeps k_apply for every dense column
    bool hp_node_stream = false;   // RTB200_HP_STREAM=node: node-parallel streaming phase of k_apply
    int64_t launches = 0;
    int sm_count = 148;

    // node-wise block-Jacobi preconditioner state (precond_circ.cuh, precond_eig.cuh)
    bool pc_ready = false;
    double pc_shift = 1.0;         // extinction floor of the preconditioner, in units of 1 / min(dx, dy)
    double* pc_eig_pool = nullptr; // per-signature eigen tables (rtb200_pc_set_eig), in allocs
    int64_t* pc_eig_off = nullptr; // [n_sig] device copy
    std::vector<int64_t> pc_eig_off_host;

    // transport-sweep / element block-Jacobi preconditioner state (precond_sweep.cu)
    int pc_type = RTB200_PC_NODE;  // what rtb200_pc_apply computes
    int pc_ncorr = -1;             // defect-correction steps of the local solves (-1: choose from the data)
    struct SweepState* sweep = nullptr;
    // host copies of the tables the sweep schedule is built from (rtb200_set_mesh)
    std::vector<int32_t> h_col_nx, h_col_ny, h_col_sig, h_col_bdry, h_face_nbr, h_sig_nt, h_th_bits;
    std::vector<int64_t> h_sig_th0, h_col_coef_off;
    std::vector<double> h_col_dx, h_col_dy, h_lgl_w, h_lgl_s;

    // halo / multi-GPU
    int rank = 0, nranks = 1;
    void* nccl_comm = nullptr;
    NcclApi* nccl = nullptr;
    int64_t n_ghost = 0, n_send = 0;
    double* ghost_dev = nullptr;
    double* send_dev = nullptr;
    int64_t* send_src_dev = nullptr;
    std::vector<int32_t> peer_rank;
    std::vector<int64_t> peer_send0, peer_recv0;
};

int rt_fail(rtb200_ctx* ctx, int code, const std::string& msg);
int rt_cuda(rtb200_ctx* ctx, cudaError_t e, const char* what);

#define RT_CUDA(ctx, call)                                        \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return rt_cuda(ctx, e__, #call);  \
    } while (0)

// apply.cu
int rt_launch_apply(rtb200_ctx* ctx, unsigned terms, int interior_only, const double* x,
                    double* y, cudaStream_t st, cudaEvent_t halo_ev = nullptr);
int rt_apply_setup(rtb200_ctx* ctx, const rtb200_flatmesh* m);
int rt_apply_with_halo(rtb200_ctx* ctx, unsigned terms, int interior_only, const double* x,
                       double* y, cudaStream_t st);
// apply_hp.cu
int rt_hp_setup(rtb200_ctx* ctx, const rtb200_flatmesh* m);
int rt_hp_launch(rtb200_ctx* ctx, const Bucket& b, unsigned terms, int interior_only, const double* x,
                 double* y, cudaStream_t st, int group0 = 0, int n_groups = -1);
// comm.cu
int rt_allreduce_sum(rtb200_ctx* ctx, double* buf_dev, int count, cudaStream_t st);
void rt_comm_destroy(rtb200_ctx* ctx);
// precond_circ.cuh / precond_eig.cuh (compiled into apply.cu)
int rt_pc_setup(rtb200_ctx* ctx, cudaStream_t st);
int rt_pc_apply(rtb200_ctx* ctx, const double* x, double* y, cudaStream_t st);
int rt_pc_apply_type(rtb200_ctx* ctx, int pc_type, const double* x, double* y, cudaStream_t st);
void rt_pc_destroy(rtb200_ctx* ctx);
// precond_sweep.cu
int rt_sweep_apply(rtb200_ctx* ctx, int pc_type, const double* r, double* z, cudaStream_t st);
void rt_sweep_destroy(rtb200_ctx* ctx);
void rt_sweep_invalidate(rtb200_ctx* ctx);      // coefficients changed
