#define PZ_STEP_AI_MASK 3
#include "pz_step_inst.inc"
