#define VIC_NN 32
#include "vicgpu_step.inc"
