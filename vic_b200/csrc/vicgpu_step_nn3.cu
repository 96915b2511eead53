#define VIC_NN 3
#include "vicgpu_step.inc"
