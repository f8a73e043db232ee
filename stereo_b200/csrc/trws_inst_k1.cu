// TRW-S kernels for up to 32 labels (1 per lane); see trws_inst.inc.
#define SB_K 1
#define SB_KOPS_NAME kops_k1
#include "trws_inst.inc"
