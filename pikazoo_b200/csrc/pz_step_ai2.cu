#define PZ_STEP_AI_MASK 2
#include "pz_step_inst.inc"
