// vicgpu_kernels.h -- launchers of the per-record kernels; one translation unit per thermal-node
// template width (vicgpu_step_nn*.cu) so that they compile in parallel.
#ifndef VICGPU_KERNELS_H
#define VICGPU_KERNELS_H
#include <cuda_runtime.h>
#include "vic_engine.cuh"

// thread block of the per-HRU kernels: default and the largest the kernels are compiled for.  384 = one block per SM at 168 registers;
// with the rows binned by kind the 12 warps of an SM then run the same code (measured: 1.15 ms vs 1.39 ms per launch at 128)
#define VICGPU_HRU_BLOCK 384
#define VICGPU_HRU_BLOCK_MAX 384

void vicgpu_launch_hru_step_nn3(const vic::Opts* d_o, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, int nsm = 0, long long sync_limit = 0);
void vicgpu_launch_hru_step_nn10(const vic::Opts* d_o, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, int nsm = 0, long long sync_limit = 0);
void vicgpu_launch_hru_step_nn32(const vic::Opts* d_o, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, int nsm = 0, long long sync_limit = 0);

void vicgpu_launch_hru_steps_nn3(const vic::Opts* d_o, const vic::Tables& t, const double* forcing, size_t per, const vic::RecBlock& rb, double* snap,
                                 size_t snap_stride, double* hdiag, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr);
void vicgpu_launch_hru_steps_nn10(const vic::Opts* d_o, const vic::Tables& t, const double* forcing, size_t per, const vic::RecBlock& rb, double* snap,
                                  size_t snap_stride, double* hdiag, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr);
void vicgpu_launch_hru_steps_nn32(const vic::Opts* d_o, const vic::Tables& t, const double* forcing, size_t per, const vic::RecBlock& rb, double* snap,
                                  size_t snap_stride, double* hdiag, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr);

#endif
