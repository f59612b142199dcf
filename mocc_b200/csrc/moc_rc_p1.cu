// register-chunk sweep kernels with 1 lane per chunk (bundles of 1 polar angle)
#define RC_P 1
#define RC_PICK pick_rc_kernel_p1
#define RC_PICK_PERSIST pick_rc_persist_kernel_p1
#include "moc_rc_inst.inc"
