// TRW-S kernels for up to 128 labels (4 per lane); see trws_inst.inc.
#define SB_K 4
#define SB_KOPS_NAME kops_k4
#include "trws_inst.inc"
