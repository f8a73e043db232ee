// Grid-native TRW-S kernels for up to 64 labels (2 per lane); see gtrws_inst.inc.
#define SB_K 2
#define SB_GOPS_NAME gops_k2
#include "gtrws_inst.inc"
