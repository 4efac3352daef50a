// C ABI of libddp.so (declared in include/ddp.h): handle lifecycle, argument validation and
// dispatch onto the sm_100a kernels.  No torch types, no exceptions across the boundary.
#include <cstdio>
#include <cstring>
#include <new>
#include "ddp_common.cuh"

namespace {

thread_local std::string g_create_err;

int fail(ddp_handle_s* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}

int cuda_fail(ddp_handle_s* h, cudaError_t e, const char* what) {
    return fail(h, DDP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(h, call)                                              \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return cuda_fail(h, e__, #call); \
    } while (0)

QPOpts qp_defaults(const ddp_boxqp_opts* o) {
    QPOpts q{100, 1e-8, 1e-8, 0.6, 1e-22, 0.1};   // boxQP.jl:29-36
    if (o && o->max_iter > 0) {
        q.max_iter = o->max_iter;
        q.min_grad = o->min_grad;
        q.min_rel_improve = o->min_rel_improve;
        q.step_dec = o->step_dec;
        q.min_step = o->min_step;
        q.armijo = o->armijo;
    }
    return q;
}

bool fill_back_params(ddp_handle_s* h, const ddp_back_pass_args* a, BackParams& P, std::string& why) {
    if (!a) { why = "args == NULL"; return false; }
    if (!a->cx.ptr || !a->cu.ptr || !a->cxx.ptr || !a->cxu.ptr || !a->cuu.ptr || !a->fx.ptr || !a->fu.ptr) {
        why = "cx, cu, cxx, cxu, cuu, fx, fu are required";
        return false;
    }
    if (!a->diverge || !a->K || !a->k || !a->Vx || !a->dV) { why = "diverge, K, k, Vx, dV outputs are required"; return false; }
    if (a->lims && !a->u.ptr) { why = "u is required when lims is given"; return false; }
    if (h->T < 1) { why = "T must be >= 1"; return false; }
    if (a->cx.stride_t == 0 && h->T > 1) { why = "cx must carry a time axis (size(cx) == (n, N), backward_pass.jl:8)"; return false; }
    if (a->cu.stride_t == 0 && h->T > 1) { why = "cu must carry a time axis (size(cu) == (m, N), backward_pass.jl:9)"; return false; }
    P.n = h->n; P.m = h->m; P.T = h->T; P.B = h->B;
    P.cx = mk(a->cx); P.cu = mk(a->cu); P.cxx = mk(a->cxx); P.cxu = mk(a->cxu); P.cuu = mk(a->cuu);
    P.fx = mk(a->fx); P.fu = mk(a->fu); P.u = mk(a->u);
    P.lambda = a->lambda; P.reg_type = a->reg_type; P.lims = a->lims; P.lims_st = a->lims_stride_t; P.active = a->active;
    if (a->lims_stride_t != 0 && a->lims_stride_t != 2 * (int64_t)h->m) { why = "lims_stride_t must be 0 or 2m"; return false; }
    P.fxx = mk(a->fxx); P.fxu = mk(a->fxu); P.fuu = mk(a->fuu);
    P.Vxx_tri = a->Vxx_tri; P.Quu_tri = a->Quu_tri; P.Quui_tri = nullptr; P.redo = nullptr; P.redo_count = nullptr;
    P.Kp = TensorD{nullptr, 0, 0}; P.kp = P.Kp; P.Sip = P.Kp; P.eta = nullptr; P.Quui = nullptr;
    P.diverge = a->diverge; P.K = a->K; P.k = a->k; P.Vx = a->Vx; P.Vxx = a->Vxx; P.Vxx1 = a->Vxx1;
    P.Quu = a->Quu; P.dV = a->dV;
    P.qp = qp_defaults(&a->qp);
    return true;
}

}  // namespace

extern "C" {

int ddp_version(void) { return DDP_VERSION; }

int ddp_device_count(void) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
    return c;
}

int ddp_create(ddp_handle_t* out, int device, int n, int m, int T, int64_t B, uint32_t flags) {
    if (!out) return fail(nullptr, DDP_ERR_INVALID, "handle pointer is NULL");
    *out = nullptr;
    if (n < 1 || n > DDP_MAX_N || m < 1 || m > DDP_MAX_M) return fail(nullptr, DDP_ERR_UNSUPPORTED, "need 1 <= n <= 64 and 1 <= m <= 16");
    if (T < 1 || B < 1) return fail(nullptr, DDP_ERR_INVALID, "need T >= 1 and B >= 1");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DDP_ERR_CUDA, "no CUDA device available: libddp has no CPU fallback");
    if (device < 0 || device >= count) return fail(nullptr, DDP_ERR_INVALID, "device index out of range");
    CU(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(nullptr, DDP_ERR_UNSUPPORTED, "libddp is built for sm_100a (Blackwell B200) only");
    ddp_handle_s* h = new (std::nothrow) ddp_handle_s();
    if (!h) return fail(nullptr, DDP_ERR_NOMEM, "out of host memory");
    h->device = device; h->n = n; h->m = m; h->T = T; h->B = B; h->flags = flags;
    h->sm_count = prop.multiProcessorCount;
    h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->launches = 0;
    h->own_stream = true;
    e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaStreamCreate"); }
    *out = h;
    return DDP_OK;
}

int ddp_destroy(ddp_handle_t h) {
    if (!h) return DDP_OK;
    cudaSetDevice(h->device);
    if (h->cache && h->cache_free) h->cache_free(h->cache);
    if (h->comm) ddp_comm_destroy(h);
    if (h->ws) cudaFree(h->ws);
    if (h->redo) cudaFree(h->redo);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return DDP_OK;
}

const char* ddp_last_error(ddp_handle_t h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int ddp_set_stream(ddp_handle_t h, void* s) {
    if (!h) return DDP_ERR_INVALID;
    if (h->own_stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
    h->stream = (cudaStream_t)s;      // NULL is the legacy default stream (what torch's default stream is)
    return DDP_OK;
}

int ddp_synchronize(ddp_handle_t h) {
    if (!h) return DDP_ERR_INVALID;
    CU(h, cudaStreamSynchronize(h->stream));
    return DDP_OK;
}

const char* ddp_kernel_variant(ddp_handle_t h) {
    if (!h) return "";
    if (h->n == 32 && h->m == 8) return "tile32x8";
    if (h->n == 4 && h->m == 1) return "small4x1";
    return "generic";
}

int64_t ddp_launch_count(ddp_handle_t h) { return h ? h->launches : 0; }

int ddp_malloc(ddp_handle_t h, void** dptr, size_t bytes) {
    if (!h || !dptr) return DDP_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(h, DDP_ERR_NOMEM, "cudaMalloc: out of device memory"); }
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc");
    return DDP_OK;
}

int ddp_free(ddp_handle_t h, void* dptr) {
    if (!h) return DDP_ERR_INVALID;
    CU(h, cudaFree(dptr));
    return DDP_OK;
}

int ddp_memset(ddp_handle_t h, void* dptr, int value, size_t bytes) {
    if (!h) return DDP_ERR_INVALID;
    CU(h, cudaMemsetAsync(dptr, value, bytes, h->stream));
    return DDP_OK;
}

int ddp_upload(ddp_handle_t h, void* dst, const void* src, size_t bytes) {
    if (!h) return DDP_ERR_INVALID;
    CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return DDP_OK;
}

int ddp_download(ddp_handle_t h, void* dst, const void* src, size_t bytes) {
    if (!h) return DDP_ERR_INVALID;
    CU(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return DDP_OK;
}

int ddp_host_alloc(void** hptr, size_t bytes) {
    if (!hptr) return DDP_ERR_INVALID;
    return cudaMallocHost(hptr, bytes) == cudaSuccess ? DDP_OK : DDP_ERR_NOMEM;
}

int ddp_host_free(void* hptr) { return cudaFreeHost(hptr) == cudaSuccess ? DDP_OK : DDP_ERR_CUDA; }

// ---------------------------------------------------------------------------------------------
int ddp_back_pass_f64(ddp_handle_t h, const ddp_back_pass_args* a) {
    if (!h) return DDP_ERR_INVALID;
    BackParams P;
    std::string why;
    if (!fill_back_params(h, a, P, why)) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_f64: " + why);
    if (!a->lambda) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_f64: lambda is required");
    if (a->reg_type != 1 && a->reg_type != 2) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_f64: reg_type must be 1 or 2");
    CU(h, cudaSetDevice(h->device));
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) {                // flag bit 0 forces the generic kernel (used by the parity tests)
        rc = launch_back_pass_tile(h, P, false, &handled);
        if (!handled && rc == 0) rc = launch_back_pass_small(h, P, false, &handled);
    }
    if (!handled && rc == 0) rc = launch_back_pass_generic(h, P, false);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "back_pass launch");
    return DDP_OK;
}

int ddp_back_pass_gps_f64(ddp_handle_t h, const ddp_back_pass_args* a, const ddp_gps_args* g) {
    if (!h) return DDP_ERR_INVALID;
    BackParams P;
    std::string why;
    if (!fill_back_params(h, a, P, why)) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_gps_f64: " + why);
    if (!g || !g->K_prev.ptr || !g->Sigi_prev.ptr || !g->eta) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_gps_f64: K_prev, Sigi_prev, eta are required");
    if (!a->Quu || !g->Quui) return fail(h, DDP_ERR_INVALID, "ddp_back_pass_gps_f64: Quu and Quui outputs are required");
    P.Kp = mk(g->K_prev); P.kp = mk(g->k_prev); P.Sip = mk(g->Sigi_prev); P.eta = g->eta; P.Quui = g->Quui; P.Quui_tri = g->Quui_tri;
    if (P.fxx.p || P.fxu.p || P.fuu.p) return fail(h, DDP_ERR_UNSUPPORTED, "ddp_back_pass_gps_f64: second-order terms are not part of back_pass_gps (backward_pass.jl:259)");
    P.reg_type = 0; P.lambda = nullptr;
    CU(h, cudaSetDevice(h->device));
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) rc = launch_back_pass_tile(h, P, true, &handled);
    if (!handled && rc == 0) rc = launch_back_pass_generic(h, P, true);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "back_pass_gps launch");
    return DDP_OK;
}

int ddp_boxqp_f64(ddp_handle_t h, int64_t B, const double* H, const double* g, const double* lower, const double* upper,
                  const double* x0, const ddp_boxqp_opts* opts, double* x, int32_t* result, double* Hfree,
                  uint32_t* free_mask, int32_t* nfactor) {
    if (!h) return DDP_ERR_INVALID;
    if (!H || !g || !lower || !upper || !x0 || !x || !result) return fail(h, DDP_ERR_INVALID, "ddp_boxqp_f64: H, g, lower, upper, x0, x, result are required");
    if (B < 0) return fail(h, DDP_ERR_INVALID, "ddp_boxqp_f64: B < 0");
    CU(h, cudaSetDevice(h->device));
    int rc = launch_boxqp(h, B, h->m, H, g, lower, upper, x0, qp_defaults(opts), x, result, Hfree, free_mask, nfactor);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "boxqp launch");
    return DDP_OK;
}

static bool fill_model(ddp_handle_s* h, const ddp_model* m, ModelD& M, std::string& why) {
    if (!m) { why = "model == NULL"; return false; }
    M.kind = m->kind;
    M.A = mk(m->A); M.Bm = mk(m->Bm); M.Q = mk(m->Q); M.R = mk(m->R); M.goal = m->goal;
    for (int i = 0; i < 8; i++) M.p[i] = m->p[i];
    M.terminal_cost = m->terminal_cost ? 1 : 0;
    M.flags = m->flags;
    if (m->kind == DDP_MODEL_LINEAR) {
        if (!m->A.ptr || !m->Bm.ptr) { why = "linear model needs A and B"; return false; }
    } else if (m->kind == DDP_MODEL_PENDCART) {
        if (h->n != 4 || h->m != 1) { why = "pendcart model needs n == 4, m == 1"; return false; }
    } else {
        why = "unknown model kind: arbitrary host callbacks cannot run on the device (no CPU fallback)";
        return false;
    }
    if (!m->Q.ptr || !m->R.ptr) { why = "model needs Q and R"; return false; }
    return true;
}

int ddp_forward_pass_f64(ddp_handle_t h, const ddp_model* model, const ddp_forward_pass_args* a) {
    if (!h) return DDP_ERR_INVALID;
    if (!a) return fail(h, DDP_ERR_INVALID, "ddp_forward_pass_f64: args == NULL");
    FwdParams P;
    std::string why;
    if (!fill_model(h, model, P.model, why)) return fail(h, model && model->kind > 2 ? DDP_ERR_UNSUPPORTED : DDP_ERR_INVALID, "ddp_forward_pass_f64: " + why);
    if (!a->x0.ptr || !a->u.ptr || !a->xnew || !a->unew || !a->cost) return fail(h, DDP_ERR_INVALID, "ddp_forward_pass_f64: x0, u, xnew, unew, cost are required");
    if ((a->K == nullptr) != (a->k == nullptr)) return fail(h, DDP_ERR_INVALID, "ddp_forward_pass_f64: K and k must both be given or both be NULL");
    if (a->K && !a->x.ptr) return fail(h, DDP_ERR_INVALID, "ddp_forward_pass_f64: x is required with a policy");
    P.n = h->n; P.m = h->m; P.T = h->T; P.B = h->B;
    P.K = a->K; P.k = a->k; P.x0 = mk(a->x0); P.x = mk(a->x); P.u = mk(a->u);
    P.alpha = a->alpha; P.alpha_scalar = a->alpha_scalar; P.u_scale = (a->u_scale == 0.0) ? 1.0 : a->u_scale;
    P.lims = a->lims; P.lims_st = a->lims_stride_t; P.active = a->active;
    if (a->lims_stride_t != 0 && a->lims_stride_t != 2 * (int64_t)h->m) return fail(h, DDP_ERR_INVALID, "ddp_forward_pass_f64: lims_stride_t must be 0 or 2m");
    P.xnew = a->xnew; P.unew = a->unew; P.cost = a->cost; P.cost_t = a->cost_t; P.cx = a->cx; P.cu = a->cu;
    CU(h, cudaSetDevice(h->device));
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) rc = launch_forward_fast(h, P, &handled);
    if (!handled && rc == 0) rc = launch_forward_generic(h, P);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "forward_pass launch");
    return DDP_OK;
}

int ddp_forward_costs_multi_f64(ddp_handle_t h, const ddp_model* model, const ddp_forward_pass_args* a, int32_t n_alpha,
                                const double* alpha, double* cost_out) {
    if (!h) return DDP_ERR_INVALID;
    if (!a || !alpha || !cost_out || n_alpha < 1) return fail(h, DDP_ERR_INVALID, "ddp_forward_costs_multi_f64: args, alpha, cost_out are required");
    FwdParams P;
    std::string why;
    if (!fill_model(h, model, P.model, why)) return fail(h, model && model->kind > 2 ? DDP_ERR_UNSUPPORTED : DDP_ERR_INVALID, "ddp_forward_costs_multi_f64: " + why);
    if (!a->K || !a->k || !a->x.ptr || !a->x0.ptr || !a->u.ptr) return fail(h, DDP_ERR_INVALID, "ddp_forward_costs_multi_f64: K, k, x0, x, u are required");
    P.n = h->n; P.m = h->m; P.T = h->T; P.B = h->B;
    P.K = a->K; P.k = a->k; P.x0 = mk(a->x0); P.x = mk(a->x); P.u = mk(a->u);
    P.alpha = nullptr; P.alpha_scalar = 1.0; P.u_scale = (a->u_scale == 0.0) ? 1.0 : a->u_scale;
    P.lims = a->lims; P.lims_st = a->lims_stride_t; P.active = a->active;
    P.xnew = a->xnew; P.unew = a->unew; P.cost = a->cost; P.cost_t = nullptr; P.cx = nullptr; P.cu = nullptr;
    CU(h, cudaSetDevice(h->device));
    bool handled = false;
    int rc = 0;
    if (!(h->flags & 1u)) rc = launch_forward_multi(h, P, n_alpha, alpha, cost_out, &handled);
    if (!handled && rc == 0) {
        if (!a->xnew || !a->unew) return fail(h, DDP_ERR_INVALID, "ddp_forward_costs_multi_f64: xnew, unew scratch are required for this shape");
        for (int i = 0; i < n_alpha && rc == 0; i++) {
            P.alpha_scalar = alpha[i];
            P.cost = cost_out + (long long)i * h->B;
            bool hd = false;
            if (!(h->flags & 1u)) rc = launch_forward_fast(h, P, &hd);
            if (!hd && rc == 0) rc = launch_forward_generic(h, P);
        }
    }
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "forward_costs_multi launch");
    return DDP_OK;
}

int ddp_batch_stats_f64(ddp_handle_t h, const double* cost_old, const double* cost_new, const double* dV, const double* alpha,
                        double alpha_scalar, const int32_t* diverge, const uint8_t* active, double* stats8) {
    if (!h || !stats8) return DDP_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    int rc = launch_batch_stats(h, h->B, cost_old, cost_new, dV, alpha, alpha_scalar, diverge, active, stats8);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "batch_stats launch");
    return DDP_OK;
}

int ddp_kl_div_f64(ddp_handle_t h, const ddp_kl_args* a) {
    if (!h) return DDP_ERR_INVALID;
    if (!a || !a->fx.ptr || !a->R1.ptr || !a->xnew || !a->xold || !a->K_new || !a->k_new || !a->Sig_new || !a->K_prev.ptr ||
        !a->Sig_prev.ptr || !a->Sigi_prev.ptr || !a->kl_mean)
        return fail(h, DDP_ERR_INVALID, "ddp_kl_div_f64: missing argument");
    KlParams P;
    P.n = h->n; P.m = h->m; P.T = h->T; P.B = h->B;
    P.fx = mk(a->fx); P.R1 = mk(a->R1); P.Kp = mk(a->K_prev); P.kp = mk(a->k_prev); P.Sp = mk(a->Sig_prev); P.Sip = mk(a->Sigi_prev);
    P.xnew = a->xnew; P.xold = a->xold; P.Kn = a->K_new; P.kn = a->k_new; P.Sn = a->Sig_new;
    P.kl_t = a->kl_t; P.kl_mean = a->kl_mean; P.active = nullptr;
    if (a->Sx_mode < 0 || a->Sx_mode > 2) return fail(h, DDP_ERR_INVALID, "ddp_kl_div_f64: Sx_mode must be 0, 1 or 2");
    if (a->Sx_mode != 0 && !a->Sx_tri) return fail(h, DDP_ERR_INVALID, "ddp_kl_div_f64: Sx_mode 1 / 2 need Sx_tri");
    if (a->Sx_count < 0 || a->Sx_count > h->B) return fail(h, DDP_ERR_INVALID, "ddp_kl_div_f64: Sx_count must be in 0..B");
    P.Sx_tri = a->Sx_tri; P.sx_mode = a->Sx_mode; P.sx_count = a->Sx_count;
    CU(h, cudaSetDevice(h->device));
    int rc = launch_kl_div(h, P);
    if (rc != 0) return cuda_fail(h, (cudaError_t)rc, "kl_div launch");
    return DDP_OK;
}

}  // extern "C"
