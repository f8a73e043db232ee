// Grid-native TRW-S kernels for up to 32 labels (1 per lane); see gtrws_inst.inc.
#define SB_K 1
#define SB_GOPS_NAME gops_k1
#include "gtrws_inst.inc"
