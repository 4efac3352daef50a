/* Minimal C host of libddp.so: one backward sweep + forward rollout for a batch of LQ problems, through the
 * same entry points the Julia shim binds with ccall (include/ddp.h).  No CUDA headers, no C++.
 *
 *   gcc -O2 -I include examples/c_host.c -L differentialdynamicprogramming.jl_b200 -lddp -lm -o c_host
 *   LD_LIBRARY_PATH=differentialdynamicprogramming.jl_b200 ./c_host
 *
 * Layout reminder: arrays are column-major per trajectory with the batch as the slowest index, i.e. a Julia
 * Array of size (n,n,B) or (n,T,B) can be passed as is. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ddp.h"

#define CK(call)                                                                          \
    do {                                                                                  \
        int rc__ = (call);                                                                \
        if (rc__ != DDP_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, ddp_last_error(h)); return 1; } \
    } while (0)

static void* up(ddp_handle_t h, const void* src, size_t bytes) {
    void* d = NULL;
    if (ddp_malloc(h, &d, bytes) != DDP_OK || ddp_upload(h, d, src, bytes) != DDP_OK) return NULL;
    return d;
}

int main(void) {
    enum { n = 32, m = 8, T = 64, B = 256 };
    ddp_handle_t h = NULL;
    if (ddp_create(&h, 0, n, m, T, B, 0) != DDP_OK) { fprintf(stderr, "ddp_create: %s\n", ddp_last_error(NULL)); return 1; }
    printf("kernel variant: %s\n", ddp_kernel_variant(h));

    /* per-trajectory LTI dynamics x+ = A x + B u (A = I + small skew part), shared cost Q = hI, R = 0.1hI */
    const double hstep = 0.01;
    double* A = calloc((size_t)B * n * n, sizeof(double));
    double* Bm = calloc((size_t)B * n * m, sizeof(double));
    double *Q = calloc(n * n, sizeof(double)), *R = calloc(m * m, sizeof(double)), *cxu = calloc(n * m, sizeof(double));
    double* u = calloc((size_t)B * T * m, sizeof(double));
    double* lam = malloc(B * sizeof(double));
    srand(1);
    for (int b = 0; b < B; b++) {
        for (int j = 0; j < n; j++)
            for (int i = 0; i < n; i++) {
                double g = hstep * ((double)rand() / RAND_MAX - 0.5);
                A[((size_t)b * n + j) * n + i] += (i == j) ? 1.0 : g;          /* column-major (i,j) */
                A[((size_t)b * n + i) * n + j] -= (i == j) ? 0.0 : g;
            }
        for (int e = 0; e < n * m; e++) Bm[(size_t)b * n * m + e] = hstep * ((double)rand() / RAND_MAX - 0.5);
        lam[b] = 1.0;
        for (int t = 0; t < T; t++)
            for (int a = 0; a < m; a++) u[((size_t)b * T + t) * m + a] = 0.1 * ((double)rand() / RAND_MAX - 0.5);
    }
    for (int i = 0; i < n; i++) Q[i * n + i] = hstep;
    for (int a = 0; a < m; a++) R[a * m + a] = 0.1 * hstep;

    void *dA = up(h, A, sizeof(double) * B * n * n), *dB = up(h, Bm, sizeof(double) * B * n * m), *dQ = up(h, Q, sizeof(double) * n * n),
         *dR = up(h, R, sizeof(double) * m * m), *dcxu = up(h, cxu, sizeof(double) * n * m), *du = up(h, u, sizeof(double) * B * T * m),
         *dlam = up(h, lam, sizeof(double) * B);
    double x0h[B * n];
    for (int e = 0; e < B * n; e++) x0h[e] = 1.0;
    void* dx0 = up(h, x0h, sizeof(x0h));
    void *dx, *dun, *dcost, *dcx, *dcu, *dK, *dk, *dVx, *ddV, *ddiv, *dxn, *dun2, *dcost2;
    CK(ddp_malloc(h, &dx, sizeof(double) * B * T * n)); CK(ddp_malloc(h, &dun, sizeof(double) * B * T * m));
    CK(ddp_malloc(h, &dcost, sizeof(double) * B)); CK(ddp_malloc(h, &dcx, sizeof(double) * B * T * n));
    CK(ddp_malloc(h, &dcu, sizeof(double) * B * T * m)); CK(ddp_malloc(h, &dK, sizeof(double) * B * T * n * m));
    CK(ddp_malloc(h, &dk, sizeof(double) * B * T * m)); CK(ddp_malloc(h, &dVx, sizeof(double) * B * T * n));
    CK(ddp_malloc(h, &ddV, sizeof(double) * B * 2)); CK(ddp_malloc(h, &ddiv, sizeof(int32_t) * B));
    CK(ddp_malloc(h, &dxn, sizeof(double) * B * T * n)); CK(ddp_malloc(h, &dun2, sizeof(double) * B * T * m));
    CK(ddp_malloc(h, &dcost2, sizeof(double) * B));

    ddp_model M;
    memset(&M, 0, sizeof(M));
    M.kind = DDP_MODEL_LINEAR;
    M.A = (ddp_tensor){dA, n * n, 0};  M.Bm = (ddp_tensor){dB, n * m, 0};
    M.Q = (ddp_tensor){dQ, 0, 0};      M.R = (ddp_tensor){dR, 0, 0};
    M.flags = DDP_MODEL_Q_DIAGONAL;

    /* initial rollout (empty policy) with the fused derivative outputs cx = Qx, cu = Ru  (iLQG.jl:181-192, 225-229) */
    ddp_forward_pass_args f0;
    memset(&f0, 0, sizeof(f0));
    f0.x0 = (ddp_tensor){dx0, n, 0};  f0.u = (ddp_tensor){du, (int64_t)T * m, m};
    f0.alpha_scalar = 1.0; f0.u_scale = 1.0;
    f0.xnew = dx; f0.unew = dun; f0.cost = dcost; f0.cx = dcx; f0.cu = dcu;
    CK(ddp_forward_pass_f64(h, &M, &f0));

    /* back_pass(cx,cu,cxx,cxu,cuu,fx,fu,lambda,regType,lims,x,u)  (backward_pass.jl:217) */
    ddp_back_pass_args a;
    memset(&a, 0, sizeof(a));
    a.cx = (ddp_tensor){dcx, (int64_t)T * n, n};  a.cu = (ddp_tensor){dcu, (int64_t)T * m, m};
    a.cxx = M.Q; a.cuu = M.R; a.cxu = (ddp_tensor){dcxu, 0, 0};
    a.fx = M.A; a.fu = M.Bm;
    a.lambda = dlam; a.reg_type = 1;
    a.diverge = ddiv; a.K = dK; a.k = dk; a.Vx = dVx; a.dV = ddV;
    CK(ddp_back_pass_f64(h, &a));

    /* forward_pass(traj_new,x0,u,x,alpha=1,...)  (forward_pass.jl:9) */
    ddp_forward_pass_args f1;
    memset(&f1, 0, sizeof(f1));
    f1.K = dK; f1.k = dk;
    f1.x0 = (ddp_tensor){dx0, n, 0};  f1.x = (ddp_tensor){dx, (int64_t)T * n, n};  f1.u = (ddp_tensor){dun, (int64_t)T * m, m};
    f1.alpha_scalar = 1.0; f1.u_scale = 1.0;
    f1.xnew = dxn; f1.unew = dun2; f1.cost = dcost2;
    CK(ddp_forward_pass_f64(h, &M, &f1));
    CK(ddp_synchronize(h));

    double c0[B], c1[B], dV[2 * B];
    int32_t div[B];
    CK(ddp_download(h, c0, dcost, sizeof(c0))); CK(ddp_download(h, c1, dcost2, sizeof(c1)));
    CK(ddp_download(h, dV, ddV, sizeof(dV))); CK(ddp_download(h, div, ddiv, sizeof(div)));
    int ndiv = 0, nacc = 0;
    double s0 = 0, s1 = 0;
    for (int b = 0; b < B; b++) {
        ndiv += div[b] > 0; s0 += c0[b]; s1 += c1[b];
        nacc += (c0[b] - c1[b]) / (-(dV[2 * b] + dV[2 * b + 1])) > 0;       /* ratio of iLQG.jl:272-276 at alpha = 1 */
    }
    printf("%d trajectories: mean cost %.6f -> %.6f, %d diverged back passes, %d accepted at alpha = 1, %lld kernel launches\n", B, s0 / B,
           s1 / B, ndiv, nacc, (long long)ddp_launch_count(h));
    ddp_destroy(h);
    return (ndiv == 0 && nacc == B) ? 0 : 2;
}
