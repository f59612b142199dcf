// register-chunk sweep kernels with 2 lanes per chunk (bundles of 2 polar angles)
#define RC_P 2
#define RC_PICK pick_rc_kernel_p2
#define RC_PICK_PERSIST pick_rc_persist_kernel_p2
#include "moc_rc_inst.inc"
