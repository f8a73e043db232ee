// TRW-S kernels for up to 64 labels (2 per lane); see trws_inst.inc.
#define SB_K 2
#define SB_KOPS_NAME kops_k2
#include "trws_inst.inc"
