// Grid-native TRW-S kernels for up to 96 labels (3 per lane); see gtrws_inst.inc.
#define SB_K 3
#define SB_GOPS_NAME gops_k3
#include "gtrws_inst.inc"
