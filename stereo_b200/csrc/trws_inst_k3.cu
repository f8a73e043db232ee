// TRW-S kernels for up to 96 labels (3 per lane); see trws_inst.inc.
#define SB_K 3
#define SB_KOPS_NAME kops_k3
#include "trws_inst.inc"
