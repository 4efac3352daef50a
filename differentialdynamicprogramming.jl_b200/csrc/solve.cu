// placeholder: device-resident iLQG driver and host-buffer iteration pipeline (filled in next)
#include "ddp_common.cuh"
extern "C" {
int ddp_ilqg_solve_f64(ddp_handle_t h, const ddp_model*, const ddp_ilqg_opts*, const double*, const double*, double*, double*, double*, double*, double*, double*, ddp_ilqg_state*, int32_t*) {
    if (h) h->err = "ddp_ilqg_solve_f64: not built yet"; return DDP_ERR_UNSUPPORTED; }
int ddp_ilqg_iter_host_f64(ddp_handle_t h, ddp_iter_host_args*) {
    if (h) h->err = "ddp_ilqg_iter_host_f64: not built yet"; return DDP_ERR_UNSUPPORTED; }
}
