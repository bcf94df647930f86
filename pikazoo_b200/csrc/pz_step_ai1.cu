#define PZ_STEP_AI_MASK 1
#include "pz_step_inst.inc"
