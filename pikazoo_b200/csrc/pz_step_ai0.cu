#define PZ_STEP_AI_MASK 0
#include "pz_step_inst.inc"
