/* ddp.h -- C ABI of libddp.so: the B200-native batched iLQG/DDP hot path.
 *
 * The reference (baggepinnen/DifferentialDynamicProgramming.jl v0.5.0) is pure Julia and has no
 * FFI of its own; the boundary it offers is its method surface.  Each entry point below replaces
 * one reference method, batched over B independent trajectories, and is what a Julia `ccall`
 * (see INTEGRATION.md / julia/) or the Python ctypes mirror binds:
 *
 *   ddp_back_pass_f64      <- back_pass(cx,cu,cxx,cxu,cuu,fx,fu,λ,regType,lims,x,u)
 *                             src/backward_pass.jl:162-252 (+ macros :3-79)
 *   ddp_back_pass_gps_f64  <- back_pass_gps(cx,cu,cxx,cxu,cuu,fx,fu,lims,x,u,kl_cost_terms)
 *                             src/backward_pass.jl:259-350 (+ ∇kl, src/klutils.jl:8-23)
 *   ddp_boxqp_f64          <- boxQP(H,g,lower,upper,x0)                 src/boxQP.jl:29-188
 *   ddp_forward_pass_f64   <- forward_pass(traj,x0,u,x,α,f,costfun,lims,diff)
 *                             src/forward_pass.jl:9-33 (f/costfun = built-in model descriptors)
 *   ddp_kl_div_f64         <- forward_covariance + kl_div_wiki
 *                             src/forward_pass.jl:37-56, src/klutils.jl:70-100
 *   ddp_ilqg_solve_f64     <- iLQG(f,costfun,df,x0,u0;kw...)           src/iLQG.jl:143-341
 *   ddp_ilqgkl_solve_f64   <- iLQGkl outer loop + calc_eta (src/iLQGkl.jl:93-183, src/klutils.jl:110-130)
 *   ddp_ilqg_iter_host_f64 <- one back_pass + forward_pass (α given) on HOST buffers, chunked
 *                             and overlapped with the PCIe copies (the end-to-end path)
 *
 * Conventions
 *   - All arithmetic is FP64.  No exceptions cross the ABI: every function returns DDP_OK (0) or
 *     a negative error code; ddp_last_error() gives the text.  Numerical outcomes (Cholesky
 *     failure, QP result codes) are per-trajectory output arrays, never errors.
 *   - Memory layout is the reference's column-major Julia arrays with the batch appended as the
 *     trailing dimension, i.e. C order [B][T][col][row]:  fx (n,n,T,B), fu (n,m,T,B),
 *     cx (n,T,B), cu (m,T,B), K (m,n,T,B), k (m,T,B), Vx (n,T,B), Vxx (n,n,T,B), Quu (m,m,T,B).
 *   - Input tensors are described by {device pointer, batch stride, time stride} in elements;
 *     a stride of 0 broadcasts.  This replaces the reference's dispatch on array rank
 *     (time-invariant 2-D vs time-varying 3-D arguments, backward_pass.jl:162/179/217).
 *   - `diverge` is the reference's 1-based timestep index of the failed step, 0 on success.
 *   - A handle owns one CUDA stream (or borrows one via ddp_set_stream) and is not thread-safe.
 *     Calls are asynchronous on that stream unless stated; ddp_synchronize() waits.
 *   - There is NO CPU fallback: without a CUDA device ddp_create fails.
 */
#ifndef DDP_H_
#define DDP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DDP_API __attribute__((visibility("default")))
#else
#define DDP_API
#endif

#define DDP_VERSION 201          /* 0.2.1 */
#define DDP_MAX_N 64
#define DDP_MAX_M 16

enum {
    DDP_OK = 0,
    DDP_ERR_INVALID = -1,        /* bad argument / shape */
    DDP_ERR_CUDA = -2,           /* CUDA runtime error */
    DDP_ERR_UNSUPPORTED = -3,    /* size or option outside the built kernels */
    DDP_ERR_NOMEM = -4,
    DDP_ERR_INCOMPLETE = -5      /* a driver loop hit its safety cap; per-trajectory status says which ones */
};

typedef struct ddp_handle_s* ddp_handle_t;

/* Strided view of a device array of doubles.  ptr == NULL means "absent". */
typedef struct ddp_tensor {
    const double* ptr;
    int64_t stride_b;            /* elements between consecutive trajectories (0 = shared)      */
    int64_t stride_t;            /* elements between consecutive timesteps    (0 = time-invariant) */
} ddp_tensor;

/* boxQP options, defaults of src/boxQP.jl:29-36 */
typedef struct ddp_boxqp_opts {
    int32_t max_iter;            /* 100   */
    double min_grad;             /* 1e-8  */
    double min_rel_improve;      /* 1e-8  */
    double step_dec;             /* 0.6   */
    double min_step;             /* 1e-22 */
    double armijo;               /* 0.1   */
} ddp_boxqp_opts;

/* ---- lifecycle ------------------------------------------------------------------------- */
DDP_API int ddp_version(void);
DDP_API int ddp_device_count(void);
/* n <= 64 states, m <= 16 controls, T timesteps (= N of the reference), B trajectories. */
DDP_API int ddp_create(ddp_handle_t* h, int device, int n, int m, int T, int64_t B, uint32_t flags);
DDP_API int ddp_destroy(ddp_handle_t h);
DDP_API const char* ddp_last_error(ddp_handle_t h);           /* h may be NULL: last create error */
DDP_API int ddp_set_stream(ddp_handle_t h, void* cuda_stream); /* borrow a cudaStream_t (NULL = default stream) */
DDP_API int ddp_synchronize(ddp_handle_t h);
/* which kernel family the dimensions of this handle dispatch to: "tile32x8", "small4x1", "generic" */
DDP_API const char* ddp_kernel_variant(ddp_handle_t h);
/* number of kernels launched through this handle since creation (bench.py's gpu_launches) */
DDP_API int64_t ddp_launch_count(ddp_handle_t h);

/* ---- device memory (so a Julia/C host needs no CUDA binding of its own) ------------------ */
DDP_API int ddp_malloc(ddp_handle_t h, void** dptr, size_t bytes);
DDP_API int ddp_free(ddp_handle_t h, void* dptr);
DDP_API int ddp_memset(ddp_handle_t h, void* dptr, int value, size_t bytes);
DDP_API int ddp_upload(ddp_handle_t h, void* dst_dev, const void* src_host, size_t bytes);    /* sync */
DDP_API int ddp_download(ddp_handle_t h, void* dst_host, const void* src_dev, size_t bytes);  /* sync */
DDP_API int ddp_host_alloc(void** hptr, size_t bytes);        /* pinned host memory */
DDP_API int ddp_host_free(void* hptr);

/* ---- backward pass ---------------------------------------------------------------------- */
typedef struct ddp_back_pass_args {
    /* inputs (device) */
    ddp_tensor cx;               /* (n,T,B)                                   */
    ddp_tensor cu;               /* (m,T,B)                                   */
    ddp_tensor cxx;              /* (n,n[,T][,B])                             */
    ddp_tensor cxu;              /* (n,m[,T][,B])  (used transposed, as the reference does) */
    ddp_tensor cuu;              /* (m,m[,T][,B])                             */
    ddp_tensor fx;               /* (n,n[,T][,B])  fx[i,j] = df_i/dx_j        */
    ddp_tensor fu;               /* (n,m[,T][,B])                             */
    const double* lambda;        /* [B] Levenberg parameter per trajectory (ignored by gps)      */
    int32_t reg_type;            /* 1: Quu + λI, 2: Vxx + λI  (iLQG.jl regType)                   */
    const double* lims;          /* (m,2) column-major [lower(m); upper(m)], NULL or lower[0] >  */
                                 /* upper[0]  => Cholesky branch (backward_pass.jl:31)           */
    int64_t lims_stride_t;       /* 0: one (m,2) block for all steps (the reference); 2m: lims is (m,2,T), time-varying  */
                                 /* limits (the branch test of :31 then reads step 1's block)               */
    ddp_tensor u;                /* (m,T,B), read only when lims is active                       */
    const uint8_t* active;       /* [B] or NULL; trajectories with active[b]==0 are skipped      */
    /* optional second-order dynamics terms of the 15-argument back_pass (backward_pass.jl:81-160, quirk Q12), any of them
     * absent (ptr == NULL) = `isempty`: fxx (n,n,n[,T][,B]), fxu (n,n,m[,T][,B]), fuu (n,m,m[,T][,B]), column-major with
     * the OUTPUT index k of f first:  Qxx[p,q] += sum_k Vx_k fxx[k,q,p],  Qux[a,j] += sum_k Vx_k fxu[k,j,a],
     * Quu[a,c] += sum_k Vx_k fuu[k,c,a]  (the contraction `vectens` of backward_pass.jl:1 intends; the regularised copies
     * Qux_reg / QuuF receive the same terms, :120-123). */
    ddp_tensor fxx, fxu, fuu;
    /* outputs (device) */
    int32_t* diverge;            /* [B] 1-based failed timestep, 0 = ok                          */
    double* K;                   /* (m,n,T,B)                                                    */
    double* k;                   /* (m,T,B)                                                      */
    double* Vx;                  /* (n,T,B)                                                      */
    double* Vxx;                 /* (n,n,T,B) or NULL: full value-Hessian history is optional    */
    double* Vxx1;                /* (n,n,B) or NULL: Vxx at the first timestep only              */
    double* Quu;                 /* (m,m,T,B) or NULL: unregularised Quu (the policy's Σi)       */
    double* dV;                  /* (2,B) expected cost reduction terms                          */
    /* packed upper-triangle histories: entry (r,c), r <= c, of step t at [c(c+1)/2 + r + tri*(t + T b)];
     * half the bytes of Vxx / Quu (Vxx is exactly symmetric, backward_pass.jl:71-72; of Quu the upper triangle is what
     * cholesky(Hermitian(.)) reads).  Independent of Vxx / Quu above; either, both or neither may be given. */
    double* Vxx_tri;             /* (n(n+1)/2,T,B) or NULL */
    double* Quu_tri;             /* (m(m+1)/2,T,B) or NULL */
    ddp_boxqp_opts qp;           /* used when lims is active; max_iter == 0 => defaults          */
} ddp_back_pass_args;

DDP_API int ddp_back_pass_f64(ddp_handle_t h, const ddp_back_pass_args* a);

/* KL-augmented sweep.  The KL cost terms of ∇kl (klutils.jl:8-23) are formed on the fly from the
 * previous policy; cxx/cxu/cuu/fx/fu may be time-invariant (stride_t = 0) as a convenience. */
typedef struct ddp_gps_args {
    ddp_tensor K_prev;           /* (m,n,T,B)                         */
    ddp_tensor k_prev;           /* (m,T,B) or absent (= zeros, as iLQGkl.jl:52 sets) */
    ddp_tensor Sigi_prev;        /* (m,m,T,B)  traj_prev.Σi           */
    const double* eta;           /* [B] dual variable η per trajectory */
    double* Quui;                /* (m,m,T,B) out: Σ = inv(Quu), required */
    double* Quui_tri;            /* (m(m+1)/2,T,B) or NULL: packed upper triangle of Σ (exactly symmetric) */
} ddp_gps_args;

/* a->Quu is required here (it is the new policy's Σi); a->lambda / a->reg_type are ignored. */
DDP_API int ddp_back_pass_gps_f64(ddp_handle_t h, const ddp_back_pass_args* a, const ddp_gps_args* g);

/* ---- box-constrained QP ------------------------------------------------------------------ */
/* H (m,m,B), g/lower/upper/x0 (m,B) -> x (m,B), result[B] (0..6 as boxQP.jl:172-179, or -1 where
 * the reference's cholesky would throw), Hfree (m,m,B; leading nfree x nfree block, upper),
 * free[B] bitmask (bit i = dimension i free), nfactor[B].  Any output may be NULL except x,result.
 * m here is the handle's m.  Bit-reproducible against oracle/ (fixed summation order, no FMA). */
DDP_API int ddp_boxqp_f64(ddp_handle_t h, int64_t B, const double* H, const double* g, const double* lower,
                  const double* upper, const double* x0, const ddp_boxqp_opts* opts, double* x,
                  int32_t* result, double* Hfree, uint32_t* free_mask, int32_t* nfactor);

/* Large problems (the reference's demoQP: n = 500, boxQP.jl:190-199): the same iteration with ONE CTA per problem, 1 <= n <= 1024
 * (n is an argument, not the handle's m).  H (n,n,B), g/lower/upper/x0 (n,B) -> x (n,B), result[B], Hfree (n,n,B; REQUIRED: it is the
 * work matrix of the factorisation and returns the nfree x nfree upper factor in its leading block, zeros elsewhere), free (n,B) bytes
 * (1 = free) or NULL, nfactor[B] or NULL.  The Cholesky factor reproduces the oracle's bit for bit (same subtraction order); sums over n
 * are tree reductions, so result codes / free sets agree with the m <= 16 path except on exact ties (see csrc/boxqp_large.cu). */
DDP_API int ddp_boxqp_large_f64(ddp_handle_t h, int32_t n, int64_t B, const double* H, const double* g, const double* lower,
                                const double* upper, const double* x0, const ddp_boxqp_opts* opts, double* x, int32_t* result,
                                double* Hfree, uint8_t* free_out, int32_t* nfactor);

/* ---- models: the reference's user callbacks f / costfun / df as device descriptors ------- */
enum { DDP_MODEL_LINEAR = 1, DDP_MODEL_PENDCART = 2 };
enum { DDP_MODEL_Q_DIAGONAL = 1 };   /* ddp_model.flags: Q is diagonal (e.g. Q = h*I of demo_linear.jl:18): only its diagonal is read */

typedef struct ddp_model {
    int32_t kind;
    /* DDP_MODEL_LINEAR (demo_linear.jl:35-50): x+ = A x + B u, cost = ½Σ x'Qx + ½Σ u'Ru        */
    ddp_tensor A;                /* (n,n[,T][,B]) */
    ddp_tensor Bm;               /* (n,m[,T][,B]) */
    /* both kinds */
    ddp_tensor Q;                /* (n,n[,B]) state cost  */
    ddp_tensor R;                /* (m,m[,B]) control cost */
    const double* goal;          /* [n] or NULL (zeros); cost is on x - goal                      */
    /* DDP_MODEL_PENDCART (system_pendcart.jl:51-54,83-106): Euler step; p = {g, l, h, d}         */
    double p[8];
    int32_t terminal_cost;       /* 1: add ½ d'Qd at the last state again (cost has T+1 entries,  */
                                 /*    system_pendcart.jl:104)                                    */
    int32_t flags;               /* DDP_MODEL_Q_DIAGONAL or 0 (a hint the host can check before upload) */
} ddp_model;

/* ---- forward pass ------------------------------------------------------------------------ */
typedef struct ddp_forward_pass_args {
    const double* K;             /* (m,n,T,B) or NULL: empty policy (iLQG.jl:185 initial rollout) */
    const double* k;             /* (m,T,B)  or NULL                                              */
    ddp_tensor x0;               /* (n[,B]) initial state                                         */
    ddp_tensor x;                /* (n,T,B) previous trajectory (absent for an empty policy)      */
    ddp_tensor u;                /* (m,T,B) previous controls                                     */
    const double* alpha;         /* [B] per-trajectory step size, or NULL to use alpha_scalar     */
    double alpha_scalar;
    double u_scale;              /* multiplies u before use (the initial rollout's αi*u); 0 => 1  */
    const double* lims;          /* (m,2) or NULL; clamp is applied whenever non-NULL (Q6)        */
    int64_t lims_stride_t;       /* 0, or 2m for time-varying limits (m,2,T)                      */
    const uint8_t* active;       /* [B] or NULL                                                   */
    /* outputs */
    double* xnew;                /* (n,T,B) */
    double* unew;                /* (m,T,B) */
    double* cost;                /* [B] total cost Σ_t                                            */
    double* cost_t;              /* (T+terminal,B) per-step cost or NULL                          */
    /* optional fused derivative outputs for the next backward pass (linear/pendcart cost):
     * cx = Q (x - goal), cu = R u  (demo_linear.jl:38-39, system_pendcart.jl:108-112)            */
    double* cx;                  /* (n,T,B) or NULL */
    double* cu;                  /* (m,T,B) or NULL */
} ddp_forward_pass_args;

DDP_API int ddp_forward_pass_f64(ddp_handle_t h, const ddp_model* model, const ddp_forward_pass_args* a);

/* Multi-alpha line-search helper (the serial backtracking of iLQG.jl:267-281 evaluated in one pass): the total
 * cost of the rollout for each of n_alpha step sizes, cost_out (n_alpha,B) row-major (row i = alpha[i], HOST array
 * of step sizes).  For the headline shape (n=32, m=8, per-trajectory LTI linear model; up to 10 step sizes per pass) and for
 * the pendulum-on-a-cart model (up to 8 per pass) the policy gains K are streamed ONCE and the costs are bit-identical to
 * ddp_forward_pass_f64's; other shapes run one rollout per step size.  a->alpha is ignored; a->xnew, a->unew, a->cost are scratch (contents undefined
 * on return); roll out the accepted step size with ddp_forward_pass_f64. */
DDP_API int ddp_forward_costs_multi_f64(ddp_handle_t h, const ddp_model* model, const ddp_forward_pass_args* a,
                                        int32_t n_alpha, const double* alpha, double* cost_out);

/* ---- derivatives of the built-in models (the reference's df callback, STEP 1 of iLQG.jl:225-229) ---- */
/* x (n,T,B), u (m,T,B) -> cx = Q(x-goal) (n,T,B), cu = R u (m,T,B); pendcart also fx (4,4,T,B), fu (4,1,T,B)
 * (ZoH Jacobians, system_pendcart.jl:137-154).  For the linear model fx/fu are A/B: pass NULL. */
DDP_API int ddp_model_derivs_f64(ddp_handle_t h, const ddp_model* model, const double* x, const double* u, double* fx,
                                 double* fu, double* cx, double* cu);

/* ---- batch statistics: the vector a multi-GPU run all-reduces once per iteration --------- */
/* stats[0]=Σcost_new, [1]=Σ(cost_old-cost_new), [2]=Σ expected reduction (α=alpha), [3]=#accepted
 * (ratio > 0), [4]=#diverged back passes, [5]=#active.  All doubles so one SUM all-reduce does. */
DDP_API int ddp_batch_stats_f64(ddp_handle_t h, const double* cost_old, const double* cost_new, const double* dV,
                        const double* alpha, double alpha_scalar, const int32_t* diverge,
                        const uint8_t* active, double* stats8 /* device, 8 doubles */);

/* ---- multi-GPU: the one collective of the path ----------------------------- */
/* Trajectories are independent, so a batch shards over GPUs with no data-path exchange; what the ranks share per
 * iteration is the 64-byte statistics vector of ddp_batch_stats_f64 (the line-search cost reduction).  These entry points
 * let a host without any CUDA/NCCL binding of its own (the Julia shim) do that all-reduce: rank 0 calls
 * ddp_comm_unique_id and hands the 128 bytes to the other ranks by whatever channel it has, every rank calls
 * ddp_comm_init, then ddp_comm_allreduce_stats_f64 sums stats8 in place (on the handle's stream) over NCCL / NVLink.
 * libnccl.so.2 is loaded at run time (dlopen); DDP_ERR_UNSUPPORTED if it is not present. */
DDP_API int ddp_comm_unique_id(void* id128 /* out: 128 bytes */);
DDP_API int ddp_comm_init(ddp_handle_t h, int32_t nranks, int32_t rank, const void* id128);
DDP_API int ddp_comm_allreduce_stats_f64(ddp_handle_t h, double* stats8 /* device, 8 doubles, in place */);
DDP_API int ddp_comm_destroy(ddp_handle_t h);

/* ---- KL divergence between the new and previous policy ----------------------------------- */
typedef struct ddp_kl_args {
    ddp_tensor fx;               /* (n,n[,T][,B]) model Jacobian used by forward_covariance       */
    ddp_tensor R1;               /* (n,n[,B]) process-noise covariance (Σ0 = R1)                  */
    const double* xnew;          /* (n,T,B) */
    const double* xold;          /* (n,T,B) */
    const double* K_new;         /* (m,n,T,B) */
    const double* k_new;         /* (m,T,B)   */
    const double* Sig_new;       /* (m,m,T,B) Σ  of the new policy (Quui)                          */
    ddp_tensor K_prev;           /* (m,n,T,B) */
    ddp_tensor k_prev;           /* (m,T,B) or absent (zeros) */
    ddp_tensor Sig_prev;         /* (m,m,T,B) */
    ddp_tensor Sigi_prev;        /* (m,m,T,B) */
    double* kl_t;                /* (T,B) per-step divergence (clipped at 0) or NULL               */
    double* kl_mean;             /* [B] mean over time                                             */
    /* Cache of the state covariance.  forward_covariance's state block, Sigma_{t+1} = fx Sigma_t fx' + R1 (forward_pass.jl:46),
     * depends on fx and R1 only -- not on the policy, the rollout or eta -- so every eta iteration of iLQGkl (iLQGkl.jl:93-183)
     * recomputes the same matrices (208 of this call's 248 tensor tiles per step).  Sx_tri: DEVICE (528,T,B) doubles, the upper
     * triangle of each Sigma_t packed by columns, or NULL.  Sx_mode 0: ignore; 1: compute as usual and also store; 2: read the
     * stored matrices instead of propagating (results are bit-identical).  Honoured by the n=32, m=8 kernel; other shapes
     * recompute in every mode.  Sx_count: the buffer holds the first Sx_count trajectories only ((528,T,Sx_count); 0 = all B):
     * the others are propagated in every call -- a partial cache for batches whose full cache does not fit in memory. */
    double* Sx_tri;
    int32_t Sx_mode;
    int32_t pad_;
    int64_t Sx_count;
} ddp_kl_args;

DDP_API int ddp_kl_div_f64(ddp_handle_t h, const ddp_kl_args* a);

/* ---- whole iLQG solve, device resident --------------------------------------------------- */
typedef struct ddp_ilqg_opts {          /* defaults: iLQG.jl:143-163 */
    int32_t n_alpha;                    /* <= 16 */
    double alpha[16];                   /* 10.^range(0,-3,11) */
    double tol_fun, tol_grad;           /* 1e-7, 1e-4 */
    int32_t max_iter;                   /* 500 */
    double lambda, dlambda, lambda_factor, lambda_max, lambda_min;   /* 1,1,1.6,1e10,1e-6 */
    int32_t reg_type;                   /* 1 */
    double reduce_ratio_min;            /* 0 */
    const double* lims;                 /* DEVICE (m,2) or NULL */
    /* pre-rolled start (iLQG.jl:193-197: size(x0) == (n,N) "and cost set accordingly"): DEVICE x_init (n,T,B) and
     * cost_init [B] (total cost of each trajectory); both NULL = roll out from x0 over the step sizes (:181-192).
     * With a pre-rolled start the x0 argument of ddp_ilqg_solve_f64 is ignored (x0 = x_init[:,1]). */
    const double* x_init;
    const double* cost_init;
    /* per-iteration trace (the reference's MVHistory keys, iLQG.jl:176-177, 257, 325-330): DEVICE buffer of
     * trace_cap x B records or NULL.  Record (it, b) holds what trajectory b's trace would hold at ITS iteration
     * `it` (1-based `iter` of the reference).  Entries never written stay as the caller initialised them. */
    struct ddp_ilqg_trace* trace;
    int32_t trace_cap;
} ddp_ilqg_opts;

typedef struct ddp_ilqg_trace {         /* one iteration of one trajectory */
    double lambda, dlambda;             /* :λ, :dλ   (values used by this iteration's backward pass)         */
    double cost;                        /* :cost     (total cost after the iteration; unchanged on a reject)  */
    double alpha;                       /* :α        (accepted step size; NaN on a rejected iteration, :313)  */
    double grad_norm;                   /* :grad_norm (:256-257)                                              */
    double improvement;                 /* :improvement = Δcost of the last tried step size (:326)            */
    double reduce_ratio;                /* :reduce_ratio (:327)                                               */
    int32_t accepted, bp_retries;       /* 1 = accepted; number of λ increases inside STEP 2 (:244-249)      */
} ddp_ilqg_trace;

/* per-trajectory solver state, one struct per trajectory (device array of B) */
typedef struct ddp_ilqg_state {
    double lambda, dlambda, cost, g_norm, last_dcost, last_alpha;
    int32_t iter;                       /* the reference's `iter` counter (1-based, at exit)       */
    int32_t accepted_iter;
    int32_t status;                     /* -1 running, 0 tol_grad, 1 tol_fun, 2 λ>λmax, 3 max_iter,
                                           4 initial rollout diverged (reference returns nothing),
                                           6 safety cap of the outer loop hit (call returns DDP_ERR_INCOMPLETE) */
    int32_t pad;
} ddp_ilqg_state;

/* x0 (n,B), u0 (m,T,B) device -> x (n,T,B), u (m,T,B), K, k, Vx, Vxx1 (may be NULL), cost_t NULL-able,
 * state[B].  Synchronous.  Returns the number of outer iterations executed in *n_outer if non-NULL. */
DDP_API int ddp_ilqg_solve_f64(ddp_handle_t h, const ddp_model* model, const ddp_ilqg_opts* opts, const double* x0,
                       const double* u0, double* x, double* u, double* K, double* k, double* Vx,
                       double* Vxx1, ddp_ilqg_state* state, int32_t* n_outer);

/* ---- whole iLQGkl solve (single-KL-constraint branch), device resident --------------------- */
/* Replaces the outer loop of iLQGkl (src/iLQGkl.jl:25-183, 238-252) with calc_eta (src/klutils.jl:110-130)
 * for a whole batch: every trajectory carries its own eta bracket, del0 and iteration count; the
 * KL-augmented backward sweep, the forward rollout (alpha = 1) and the KL evaluation run over the batch
 * with activity masks.  The derivatives are computed once, before the loop (iLQGkl.jl:88, quirk Q9). */
typedef struct ddp_ilqgkl_opts {        /* defaults: iLQGkl.jl:25-42 */
    double kl_step;                     /* 1.0; <= 0 => satisfied at once (klutils.jl:111)            */
    int32_t max_iter;                   /* 50                                                         */
    double eta_bracket[3];              /* {1e-8, 1, 1e16}                                            */
    double del0;                        /* 1e-4                                                       */
    int32_t max_eta_retries;            /* bounds the reference's unbounded eta-retry loop (iLQGkl.jl:97); 0 => 200 */
    const double* lims;                 /* DEVICE (m,2) or NULL                                       */
    int32_t no_covariance_cache;        /* 0: keep the state covariances of forward_covariance (they depend on fx_model and R1
                                         * only) from the first eta iteration for the later ones, for as many leading trajectories
                                         * as the free device memory holds (n=32, m=8: 4.2 KB per step and trajectory);
                                         * 1: recompute every iteration */
    int32_t pad_;
} ddp_ilqgkl_opts;

typedef struct ddp_ilqgkl_state {
    double eta_min, eta, eta_max;       /* the bracket at exit                                        */
    double del0, divergence, dcost, expected, cost;
    int32_t iter;                       /* iterations executed (1-based, as the reference's `iter`)   */
    int32_t status;                     /* -1 running, 0 KL constraint satisfied, 1 eta > 0.999 eta_max (iLQGkl.jl:178),
                                           3 max_iter, 5 eta-retry limit                              */
    int32_t retries, pad;
} ddp_ilqgkl_state;

typedef struct ddp_ilqgkl_args {
    /* inputs (device) */
    const double* x;                    /* (n,T,B) pre-rolled trajectory (iLQGkl.jl:63-70)            */
    const double* u;                    /* (m,T,B) = traj_prev.k (iLQGkl.jl:47)                       */
    const double* cost;                 /* [B] total cost of (x,u)                                    */
    ddp_tensor K_prev;                  /* (m,n,T,B) traj_prev.K                                      */
    ddp_tensor Sig_prev;                /* (m,m,T,B) traj_prev.Σ                                      */
    ddp_tensor Sigi_prev;               /* (m,m,T,B) traj_prev.Σi                                     */
    ddp_tensor fx_model;                /* (n,n[,T][,B]) what df(model,x,u) returns (forward_pass.jl:38) */
    ddp_tensor R1;                      /* (n,n[,B])     what covariance(model,x,u) returns (forward_pass.jl:42) */
    /* outputs (device) */
    double *xnew, *unew;                /* (n,T,B), (m,T,B)                                           */
    double *K, *k;                      /* new policy; k = unew on return (quirk Q11, iLQGkl.jl:241)  */
    double *Sig, *Sigi;                 /* (m,m,T,B) new policy Σ = inv(Quu), Σi = Quu                */
    double* Vx;                         /* (n,T,B)                                                    */
    double* Vxx1;                       /* (n,n,B) or NULL                                            */
    double* costnew;                    /* [B]                                                        */
    ddp_ilqgkl_state* state;            /* [B]                                                        */
} ddp_ilqgkl_args;

DDP_API int ddp_ilqgkl_solve_f64(ddp_handle_t h, const ddp_model* model, const ddp_ilqgkl_opts* opts,
                                 const ddp_ilqgkl_args* a, int32_t* n_outer);

/* ---- end-to-end iteration on HOST buffers ------------------------------------------------ */
/* One backward sweep + one forward rollout for a linear model with per-trajectory LTI dynamics
 * (the headline workload), taking pinned HOST arrays: the batch is cut into chunks whose H2D
 * copies, kernels and D2H copies are pipelined on separate streams.  The policy (K,k) stays on the
 * device (the shim's GaussianPolicy holds it and downloads lazily).  Synchronous.
 * Host inputs:  fx (n,n,B), fu (n,m,B), x (n,T,B), u (m,T,B), lambda[B], and cx (n,T,B), cu (m,T,B) -- or cx = cu = NULL,
 *               in which case cx = Q x, cu = R u are formed on the device (STEP 1 of iLQG.jl:225-229 for this model).
 * Host outputs: xnew (n,T,B), unew (m,T,B), cost[B], dV (2,B), diverge[B].
 * Q (n,n), R (m,m), cxu (n,m) are small shared host matrices.                                     */
typedef struct ddp_iter_host_args {
    const double *fx, *fu, *cx, *cu, *x, *u, *lambda;
    const double *Q, *R, *cxu;
    int32_t reg_type;
    int32_t q_diagonal;          /* 1: Q is diagonal (hint, as DDP_MODEL_Q_DIAGONAL) */
    double alpha;
    double *xnew, *unew, *cost, *dV;
    int32_t* diverge;
    int64_t chunk;               /* trajectories per chunk, 0 = library default */
    int64_t h2d_bytes, d2h_bytes; /* out: bytes moved */
    int32_t keep_policy;         /* 1: the whole batch's K,k,Vx stay on the device after the call (ddp_iter_host_policy);
                                    0: they live in one chunk-sized scratch and are gone when the call returns */
    int32_t inputs_resident;     /* 1: fx,fu,x,u,lambda are NOT uploaded: the device copies of the previous call on this
                                    handle are used (a real iLQG loop keeps them there); host pointers may be NULL */
    int32_t commit_accepted;     /* 1: after the call the device copies x,u are replaced by xnew,unew for trajectories whose
                                    step was accepted (ratio > 0, iLQG.jl:269-303), ready for inputs_resident = 1        */
    int32_t pad_;
    const double* cost_prev;     /* [B] HOST: total cost of the current (x,u); required with commit_accepted           */
} ddp_iter_host_args;

DDP_API int ddp_ilqg_iter_host_f64(ddp_handle_t h, ddp_iter_host_args* a);
/* Device pointers of what the last ddp_ilqg_iter_host_f64 call on this handle left on the device: the policy K (m,n,T,B),
 * k (m,T,B), Vx (n,T,B) (NULL unless keep_policy was set) and the rollout xnew (n,T,B), unew (m,T,B).  Owned by the handle,
 * valid until the next host iteration with another configuration or ddp_destroy.  Any out pointer may be NULL. */
DDP_API int ddp_iter_host_policy(ddp_handle_t h, double** K, double** k, double** Vx, double** xnew, double** unew);

/* ---- one iteration, device resident, chunked over a batch larger than the policy's HBM residency ------------------------ */
/* One backward sweep + one forward rollout over the handle's B trajectories for a built-in model, all arrays on the DEVICE.
 * The batch is processed in chunks of `chunk` trajectories so that the policy (K,k,Vx: 592 KiB per trajectory at n=32, m=8,
 * T=256 -- 137 GB+ at BASELINE config 5's 262144 trajectories per GPU) only ever exists for one chunk: the derivative step
 * (cx = Q(x-goal), cu = R u; pendcart also fx,fu), the backward sweep and the rollout of a chunk run back to back on the
 * handle's stream re-using one chunk-sized scratch (SURVEY section 7 "HBM capacity").  Pass K,k,Vx (full size) to keep the
 * policy instead.  Asynchronous on the handle's stream. */
typedef struct ddp_iter_args {
    const double *x, *u;         /* (n,T,B), (m,T,B) current trajectory                                   */
    const double* lambda;        /* [B]                                                                    */
    const double* alpha;         /* [B] or NULL => alpha_scalar                                            */
    double alpha_scalar;
    int32_t reg_type;            /* 1 or 2                                                                 */
    int32_t pad_;
    const double* lims;          /* (m,2) or NULL                                                          */
    const uint8_t* active;       /* [B] or NULL                                                            */
    double *xnew, *unew;         /* (n,T,B), (m,T,B)                                                       */
    double *cost, *dV;           /* [B], (2,B)                                                             */
    int32_t* diverge;            /* [B]                                                                    */
    double *K, *k, *Vx;          /* all three full size (.,T,B) to keep the policy, or all NULL (chunk scratch) */
    int64_t chunk;               /* trajectories per chunk; 0 = library default (sm_count * 8 * 55 rounds)  */
    int64_t n_chunks;            /* out */
} ddp_iter_args;

DDP_API int ddp_ilqg_iter_f64(ddp_handle_t h, const ddp_model* model, ddp_iter_args* a);

/* ---- self test: the FP64 denominators of the roofline, measured on this device ------------------------------------------ */
/* kind 0: DFMA (fma.rn.f64, 8 independent chains per thread); kind 1: DMMA (mma.sync.m8n8k4.f64, the instruction of the
 * n=32, m=8 sweep kernels); kind 2: 8 DMMA + 32 DFMA per warp-iteration interleaved (all flops counted: a rate at the DMMA
 * peak means the two instruction classes share one FP64 datapath).  Runs `reps` timed launches on the handle's stream (after one warm-up) and returns the best
 * TFLOP/s and its launch time.  Synchronous.  bench.py calls this before its timed region ("measure it"). */
DDP_API int ddp_selftest_peak_f64(ddp_handle_t h, int32_t kind, int32_t reps, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* DDP_H_ */
