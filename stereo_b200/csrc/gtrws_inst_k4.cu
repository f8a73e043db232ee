// Grid-native TRW-S kernels for up to 128 labels (4 per lane); see gtrws_inst.inc.
#define SB_K 4
#define SB_GOPS_NAME gops_k4
#include "gtrws_inst.inc"
