// TRW-S kernels for up to 256 labels (8 per lane); see trws_inst.inc.
#define SB_K 8
#define SB_KOPS_NAME kops_k8
#include "trws_inst.inc"
