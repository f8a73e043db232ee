// TRW-S kernels for up to 192 labels (6 per lane); see trws_inst.inc.
#define SB_K 6
#define SB_KOPS_NAME kops_k6
#include "trws_inst.inc"
