// Grid-native TRW-S kernels for up to 256 labels (8 per lane); see gtrws_inst.inc.
#define SB_K 8
#define SB_GOPS_NAME gops_k8
#include "gtrws_inst.inc"
