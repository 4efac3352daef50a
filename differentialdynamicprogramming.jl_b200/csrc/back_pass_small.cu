// placeholder: specialised small-n thread-per-trajectory kernel (filled in next)
#include "ddp_common.cuh"
int launch_back_pass_small(ddp_handle_s*, const BackParams&, bool, bool* handled) { *handled = false; return 0; }
