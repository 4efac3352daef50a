// Shared declarations for the libddp.so kernels (sm_100a).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/ddp.h"

struct TensorD {                 // device-side mirror of ddp_tensor
    const double* p;
    long long sb, st;
};
__host__ __device__ inline TensorD mk(const ddp_tensor& t) { return TensorD{t.ptr, (long long)t.stride_b, (long long)t.stride_t}; }
__device__ __forceinline__ const double* tp(const TensorD& t, long long b, int i) { return t.p + b * t.sb + (long long)i * t.st; }

// Julia's clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x)): NaN stays NaN (CUDA's fmin/fmax would drop it) and
// an inverted interval (lo > hi) resolves the way the reference's does (forward_pass.jl:23, boxQP.jl:58).
__device__ __forceinline__ double clamp_jl(double v, double lo, double hi) { return (v > hi) ? hi : ((v < lo) ? lo : v); }

struct QPOpts {
    int max_iter;
    double min_grad, min_rel_improve, step_dec, min_step, armijo;
};

struct BackParams {
    int n, m, T;
    long long B;
    TensorD cx, cu, cxx, cxu, cuu, fx, fu, u;
    const double* lambda;
    int reg_type;
    const double* lims;          // (m,2[,T]) or nullptr (Cholesky branch)
    long long lims_st;           // 0 or 2m (time-varying limits)
    const unsigned char* active;
    TensorD fxx, fxu, fuu;       // optional second-order dynamics terms (generic kernel only)
    // gps
    TensorD Kp, kp, Sip;
    const double* eta;
    double* Quui;
    // outputs
    int* diverge;
    double *K, *k, *Vx, *Vxx, *Vxx1, *Quu, *dV;
    double *Vxx_tri, *Quu_tri, *Quui_tri;     // packed upper-triangle histories or nullptr
    // hand-over from the tile kernel to the generic one: trajectories whose terminal cxx is not exactly symmetric (the tile
    // kernel's products assume a symmetric Vxx) are flagged in redo[] and counted in *redo_count; the generic launch that
    // follows processes only those (and returns at once when the count is zero)
    unsigned char* redo;
    int* redo_count;
    QPOpts qp;
};

struct ModelD {
    int kind;
    TensorD A, Bm, Q, R;
    const double* goal;
    double p[8];
    int terminal_cost;
    int flags;                   // DDP_MODEL_Q_DIAGONAL
};

struct FwdParams {
    int n, m, T;
    long long B;
    ModelD model;
    const double *K, *k;
    TensorD x0, x, u;
    const double* alpha;
    double alpha_scalar, u_scale;
    const double* lims;
    long long lims_st;
    const unsigned char* active;
    double *xnew, *unew, *cost, *cost_t, *cx, *cu;
};

struct KlParams {
    int n, m, T;
    long long B;
    TensorD fx, R1, Kp, kp, Sp, Sip;
    const double *xnew, *xold, *Kn, *kn, *Sn;
    double *kl_t, *kl_mean;
    const unsigned char* active;     // [B] or nullptr
    double* Sx_tri = nullptr;        // (B,T,528) packed upper triangles of the state covariances, see ddp_kl_args
    int sx_mode = 0;                 // 0 none, 1 store, 2 load
    long long sx_count = 0;          // trajectories the cache holds (0 = all)
    long long b_begin = 0, b_end = -1;   // trajectory range of one launch (kl_tile.cu; -1 = B)
};

struct ddp_handle_s {
    int device, n, m, T;
    long long B;
    uint32_t flags;
    cudaStream_t stream;
    bool own_stream;
    int sm_count;
    int max_smem_optin;
    long long launches;
    std::string err;
    void* cache = nullptr;               // lazily created pipeline state (solve.cu)
    void (*cache_free)(void*) = nullptr;
    void* comm = nullptr;                // ncclComm_t (comm.cu)
    void* ws = nullptr;                  // workspace arena of the solve drivers, kept between calls (solve.cu)
    size_t ws_cap = 0;
    unsigned char* redo = nullptr;       // tile -> generic hand-over mask (back_pass_tile.cu), redo_cap bytes + one int counter in front
    long long redo_cap = 0;
    // scratch for the solve driver / host-iteration pipeline is allocated lazily by those entry points
};

// launchers implemented in the .cu files; each returns a cudaError_t-compatible int (0 = ok) and
// sets *launched to the number of kernels it enqueued.
int launch_back_pass_generic(ddp_handle_s* h, const BackParams& P, bool gps);
int prepare_redo(ddp_handle_s* h, BackParams& P);
int launch_back_pass_tile(ddp_handle_s* h, const BackParams& P, bool gps, bool* handled);
int launch_back_pass_small(ddp_handle_s* h, const BackParams& P, bool gps, bool* handled);
int launch_forward_generic(ddp_handle_s* h, const FwdParams& P);
int launch_forward_fast(ddp_handle_s* h, const FwdParams& P, bool* handled);
int launch_forward_multi(ddp_handle_s* h, const FwdParams& P, int na, const double* alpha, double* cost_out, bool* handled);
int launch_forward_multi_tile(ddp_handle_s* h, const FwdParams& P, int na, const double* alpha, double* cost_out, bool* handled);
int launch_boxqp(ddp_handle_s* h, long long B, int m, const double* H, const double* g, const double* lower,
                 const double* upper, const double* x0, QPOpts o, double* x, int* result, double* Hfree,
                 unsigned* free_mask, int* nfactor);
int launch_kl_div(ddp_handle_s* h, const KlParams& P);
int launch_kl_div_tile(ddp_handle_s* h, const KlParams& P, bool* handled);
int launch_batch_stats(ddp_handle_s* h, long long B, const double* cost_old, const double* cost_new, const double* dV,
                       const double* alpha, double alpha_scalar, const int* diverge, const unsigned char* active,
                       double* stats8);
