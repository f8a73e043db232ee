// Grid-native TRW-S kernels for up to 192 labels (6 per lane); see gtrws_inst.inc.
#define SB_K 6
#define SB_GOPS_NAME gops_k6
#include "gtrws_inst.inc"
