// register-chunk sweep kernels with 4 lanes per chunk (bundles of 4 polar angles)
#define RC_P 4
#define RC_PICK pick_rc_kernel_p4
#define RC_PICK_PERSIST pick_rc_persist_kernel_p4
#include "moc_rc_inst.inc"
