// vicgpu_kernels.h -- launchers of the per-record kernels; one translation unit per thermal-node
// template width (vicgpu_step_nn*.cu) so that they compile in parallel.
#ifndef VICGPU_KERNELS_H
#define VICGPU_KERNELS_H
#include <cuda_runtime.h>
#include "vic_engine.cuh"

// Thread block and register budget of the per-HRU step kernel.  The kernel is latency bound, so what counts is
// how many warps an SM can interleave: at 128 registers an SM holds 16 warps.  Measured on the 10,000-cell workload (us per launch,
// winter / spring / summer / autumn week): 160 registers x 384 threads 890 / 935 / 808 / --; 128 x 448: 823 / 908 / 672 / 942;
// 128 x 512: 847 / 943 / 681 / 991; 96 x 640: 896 / -- / 766; 80 x 768: 951 / -- / 821; 64 x 1024: 996 / -- / 877.  The default block is 448
// threads (14 warps, leaving room for two one-warp blocks of the cell-output kernel on the same SM) while that covers the domain in
// one wave, else 512.  Blocks this large also keep the warps of one kind (binned rows) on one SM, which is what the instruction
// cache needs.
#ifndef VICGPU_HRU_BLOCK_MAX
#define VICGPU_HRU_BLOCK_MAX 512
#endif
#ifndef VICGPU_STEP_MAXNREG
#define VICGPU_STEP_MAXNREG 128
#endif
#ifndef VICGPU_HRU_BLOCK
#define VICGPU_HRU_BLOCK 448
#endif

// one: the configuration's model step is a single sub-step (NF == 1)
void vicgpu_launch_hru_step_nn3(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0);
void vicgpu_launch_hru_step_nn10(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0);
void vicgpu_launch_hru_step_nn32(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0);

#endif
