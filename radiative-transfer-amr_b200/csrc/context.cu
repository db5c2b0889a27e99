// Context life cycle and mesh-table upload for librtb200.so.
//
// rtb200_set_mesh copies the host tables of rtb200_flatmesh (built by
// rt_b200.dg.flatmesh from the reference's Mesh object) to the device and
// prepares the launch buckets.  This replaces the role of the five
// calc_col_matrix assemblers of the reference (src/rt/*_matrix.py): instead of
// an explicit AIJ matrix the device holds the element tables the matrix
// entries are products of.
#include "rtb200_internal.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int rt_fail(rtb200_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

int rt_cuda(rtb200_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) ctx->err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return RTB200_ECUDA;
}

static const char* k_no_ctx = "null context";

extern "C" int rtb200_abi_version(void) { return RTB200_ABI_VERSION; }

extern "C" int rtb200_create(int device, rtb200_ctx** out) {
    if (!out) return RTB200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) return RTB200_ECUDA;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return RTB200_ECUDA;
    rtb200_ctx* ctx = new (std::nothrow) rtb200_ctx();
    if (!ctx) return RTB200_EINVAL;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return RTB200_OK;
}

static void free_tables(rtb200_ctx* ctx) {
    for (void* p : ctx->allocs) cudaFree(p);
    ctx->allocs.clear();
    ctx->halo_flag_dev = nullptr;      // lives in allocs
    ctx->halo_resume_dev = nullptr;
    ctx->sched_slots.clear();          // the counters live in allocs
    ctx->buckets.clear();
    ctx->coef_dev = nullptr;
    ctx->ghost_dev = nullptr;
    ctx->send_dev = nullptr;
    ctx->send_src_dev = nullptr;
    ctx->have_mesh = false;
    rt_pc_destroy(ctx);
    rt_sweep_destroy(ctx);
}

extern "C" void rtb200_destroy(rtb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_tables(ctx);
    for (cudaStream_t q : ctx->side_streams) cudaStreamDestroy(q);
    for (cudaEvent_t e : ctx->ev_join) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_halo) cudaEventDestroy(ctx->ev_halo);
    if (ctx->ev_packed) cudaEventDestroy(ctx->ev_packed);
    ctx->ev_packed = nullptr;
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    ctx->ev_halo = nullptr;
    ctx->comm_stream = nullptr;
    ctx->side_streams.clear();
    ctx->ev_join.clear();
    ctx->ev_fork = nullptr;
    rt_comm_destroy(ctx);
    delete ctx;
}

extern "C" const char* rtb200_last_error(const rtb200_ctx* ctx) {
    return ctx ? ctx->err.c_str() : k_no_ctx;
}

extern "C" int64_t rtb200_launch_count(const rtb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

template <typename T>
static int upload(rtb200_ctx* ctx, const T* host, int64_t n, const T** dev, const char* name) {
    *dev = nullptr;
    if (n < 0) return rt_fail(ctx, RTB200_EINVAL, std::string("negative size for ") + name);
    if (n > 0 && !host) return rt_fail(ctx, RTB200_EINVAL, std::string("null table ") + name);
    int64_t alloc_n = n > 0 ? n : 1;
    T* p = nullptr;
    cudaError_t e = cudaMalloc((void**)&p, alloc_n * sizeof(T));
    if (e != cudaSuccess) return rt_cuda(ctx, e, name);
    ctx->allocs.push_back(p);
    if (n > 0) {
        e = cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return rt_cuda(ctx, e, name);
    }
    *dev = p;
    return RTB200_OK;
}

#define UP(field, n)                                                              \
    do {                                                                          \
        int rc__ = upload(ctx, m->field, (int64_t)(n), &ctx->dm.field, #field);   \
        if (rc__) return rc__;                                                    \
    } while (0)

extern "C" int rtb200_set_mesh(rtb200_ctx* ctx, const rtb200_flatmesh* m) {
    if (!ctx) return RTB200_EINVAL;
    if (!m) return rt_fail(ctx, RTB200_EINVAL, "null flatmesh");
    if (m->abi_version != RTB200_ABI_VERSION)
        return rt_fail(ctx, RTB200_EINVAL, "flatmesh ABI version mismatch");
    if (m->n_col < 0 || m->n_sig <= 0 || m->n_slab < 0 || (m->n_slab & 1))
        return rt_fail(ctx, RTB200_EINVAL, "inconsistent flatmesh sizes");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    free_tables(ctx);
    ctx->dm = DevMesh{};
    ctx->dm.n_col = m->n_col;
    ctx->dm.n_dof = m->n_dof;
    ctx->dm.n_slab = m->n_slab;
    ctx->dm.n_theta_pool = m->n_theta_pool;

    // host-side validation of the sizes the kernels assume
    for (int c = 0; c < m->n_col; ++c) {
        if (m->col_nx[c] < 1 || m->col_nx[c] > RTB200_MAX_NXY || m->col_ny[c] < 1 ||
            m->col_ny[c] > RTB200_MAX_NXY)
            return rt_fail(ctx, RTB200_EUNSUPPORTED, "nx/ny outside [1, RTB200_MAX_NXY]");
        if (m->col_sig[c] < 0 || m->col_sig[c] >= m->n_sig)
            return rt_fail(ctx, RTB200_EINVAL, "column signature id out of range");
        if (m->col_slab_off[c] & 1) return rt_fail(ctx, RTB200_EINVAL, "odd slab offset");
    }
    for (int i = 0; i < m->n_sigcell; ++i)
        if (m->cell_nth[i] < 1 || m->cell_nth[i] > RTB200_MAX_NTH)
            return rt_fail(ctx, RTB200_EUNSUPPORTED, "n_theta outside [1, RTB200_MAX_NTH]");

    const int64_t nc = m->n_col;
    UP(col_nx, nc); UP(col_ny, nc); UP(col_sig, nc); UP(col_bdry, nc);
    UP(col_dof_off, nc); UP(col_slab_off, nc); UP(col_coef_off, nc);
    UP(col_dx, nc); UP(col_dy, nc);
    UP(face_base, nc * 8); UP(face_rstride, nc * 8); UP(face_exy, nc * 8);
    UP(face_n1, nc * 8); UP(face_acpl, nc * 8);
    UP(sig_nt, m->n_sig); UP(sig_ntp, m->n_sig); UP(sig_ncell, m->n_sig);
    UP(sig_cell0, m->n_sig); UP(sig_th0, m->n_sig); UP(sig_g_off, m->n_sig);
    UP(cell_nth, m->n_sigcell); UP(cell_thoff, m->n_sigcell);
    UP(th_wt, m->n_theta_pool); UP(th_ct, m->n_theta_pool); UP(th_st, m->n_theta_pool);
    UP(th_wF, 4 * m->n_theta_pool);
    UP(th_cell, m->n_theta_pool); UP(th_bits, m->n_theta_pool);
    UP(g_pool, m->n_g_pool); UP(exy_pool, m->n_exy_pool); UP(eth_pool, m->n_eth_pool);
    UP(acpl_pool, m->n_acpl_pool);
    UP(coef, m->n_coef);
    if (!m->face_nbr && nc > 0) return rt_fail(ctx, RTB200_EINVAL, "null table face_nbr");
    if (!m->lgl_w || !m->lgl_s) return rt_fail(ctx, RTB200_EINVAL, "missing LGL tables");
    UP(face_nbr, nc * 8);
    // host copies for the sweep schedule (precond_sweep.cu)
    ctx->h_col_nx.assign(m->col_nx, m->col_nx + nc);
    ctx->h_col_ny.assign(m->col_ny, m->col_ny + nc);
    ctx->h_col_sig.assign(m->col_sig, m->col_sig + nc);
    ctx->h_col_bdry.assign(m->col_bdry, m->col_bdry + nc);
    ctx->h_face_nbr.assign(m->face_nbr, m->face_nbr + nc * 8);
    ctx->h_col_coef_off.assign(m->col_coef_off, m->col_coef_off + nc);
    ctx->h_col_dx.assign(m->col_dx, m->col_dx + nc);
    ctx->h_col_dy.assign(m->col_dy, m->col_dy + nc);
    ctx->h_sig_nt.assign(m->sig_nt, m->sig_nt + m->n_sig);
    ctx->h_sig_th0.assign(m->sig_th0, m->sig_th0 + m->n_sig);
    ctx->h_th_bits.assign(m->th_bits, m->th_bits + m->n_theta_pool);
    ctx->h_lgl_w.assign(m->lgl_w, m->lgl_w + (RTB200_MAX_NXY + 1) * RTB200_MAX_NXY);
    ctx->h_lgl_s.assign(m->lgl_s, m->lgl_s + (RTB200_MAX_NXY + 1) * RTB200_MAX_NXY * RTB200_MAX_NXY);
    ctx->coef_dev = const_cast<double*>(ctx->dm.coef);
    ctx->n_coef = m->n_coef;
    ctx->circ_pool_dev = nullptr;
    if (m->n_circ_pool > 0) {
        if (!m->sig_circ_off || !m->sig_circ_nc || !m->sig_circ_p)
            return rt_fail(ctx, RTB200_EINVAL, "circulant pool without per-signature tables");
        for (int s = 0; s < m->n_sig; ++s)
            if (m->sig_circ_off[s] >= 0 &&
                ((m->sig_circ_off[s] & 1) || m->sig_circ_nc[s] * m->sig_circ_p[s] != m->sig_nt[s] ||
                 m->sig_circ_off[s] + (int64_t)m->sig_circ_nc[s] * m->sig_circ_p[s] * m->sig_circ_p[s] >
                     m->n_circ_pool))
                return rt_fail(ctx, RTB200_EINVAL, "inconsistent circulant table of a signature");
        int rc = upload(ctx, m->circ_pool, m->n_circ_pool, &ctx->circ_pool_dev, "circ_pool");
        if (rc) return rc;
    }
    {
        const char* env = getenv("RTB200_DENSE_SCATTERING");
        ctx->use_circ = !(env && env[0] == '1');
        const char* cv = getenv("RTB200_CIRC_VERSION");
        if (cv && cv[0] == '6') ctx->circ_version = 6;
        const char* rsv = getenv("RTB200_HALO_SM_RESERVE");
        if (rsv && rsv[0] >= '0' && rsv[0] <= '9') ctx->halo_sm_reserve = atoi(rsv);
        const char* hs = getenv("RTB200_HP_STREAM");
        ctx->hp_node_stream = hs && strcmp(hs, "node") == 0;
        const char* hk = getenv("RTB200_HP_KERNEL");
        ctx->use_hp = !(hk && strcmp(hk, "old") == 0);
    }

    // halo plan
    ctx->n_ghost = m->n_ghost;
    ctx->n_send = m->n_send;
    ctx->peer_rank.assign(m->peer_rank, m->peer_rank + (m->n_peer > 0 ? m->n_peer : 0));
    ctx->peer_send0.clear();
    ctx->peer_recv0.clear();
    if (m->n_peer > 0) {
        ctx->peer_send0.assign(m->peer_send0, m->peer_send0 + m->n_peer + 1);
        ctx->peer_recv0.assign(m->peer_recv0, m->peer_recv0 + m->n_peer + 1);
    }
    {
        double* g = nullptr;
        RT_CUDA(ctx, cudaMalloc((void**)&g, (size_t)(m->n_ghost > 0 ? m->n_ghost : 2) * sizeof(double)));
        ctx->allocs.push_back(g);
        RT_CUDA(ctx, cudaMemset(g, 0, (size_t)(m->n_ghost > 0 ? m->n_ghost : 2) * sizeof(double)));
        ctx->ghost_dev = g;
        double* s = nullptr;
        RT_CUDA(ctx, cudaMalloc((void**)&s, (size_t)(m->n_send > 0 ? m->n_send : 2) * sizeof(double)));
        ctx->allocs.push_back(s);
        ctx->send_dev = s;
        const int64_t* src = nullptr;
        int rc = upload(ctx, m->send_src, m->n_send, &src, "send_src");
        if (rc) return rc;
        ctx->send_src_dev = const_cast<int64_t*>(src);
    }

    int rc = rt_apply_setup(ctx, m);
    if (rc) return rc;
    ctx->have_mesh = true;
    ctx->err.clear();
    return RTB200_OK;
}

extern "C" int rtb200_set_coeffs(rtb200_ctx* ctx, const double* coef_host, int64_t n_coef) {
    if (!ctx) return RTB200_EINVAL;
    if (!ctx->have_mesh) return rt_fail(ctx, RTB200_ESTATE, "rtb200_set_coeffs before rtb200_set_mesh");
    if (!coef_host || n_coef != ctx->n_coef)
        return rt_fail(ctx, RTB200_EINVAL, "coefficient array does not match the mesh");
    RT_CUDA(ctx, cudaSetDevice(ctx->device));
    RT_CUDA(ctx, cudaMemcpy(ctx->coef_dev, coef_host, (size_t)n_coef * sizeof(double),
                            cudaMemcpyHostToDevice));
    ctx->pc_ready = false;
    rt_sweep_invalidate(ctx);
    return RTB200_OK;
}
