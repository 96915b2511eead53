#define VIC_NN 10
#include "vicgpu_step.inc"
