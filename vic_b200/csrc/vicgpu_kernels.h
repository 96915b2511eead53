// vicgpu_kernels.h -- launchers of the per-record kernels; one translation unit per thermal-node
// template width (vicgpu_step_nn*.cu) so that they compile in parallel.
#ifndef VICGPU_KERNELS_H
#define VICGPU_KERNELS_H
#include <cuda_runtime.h>
#include "vic_engine.cuh"

// Thread block and register budget of the per-HRU step kernel.  The kernel is latency bound, so what counts is
// how many warps an SM can interleave: at 128 registers an SM holds 16 warps.  Measured on the 10,000-cell workload in round 1 (us per
// launch, winter / spring / summer / autumn week): 160 registers x 384 threads 890 / 935 / 808 / --; 128 x 448: 823 / 908 / 672 / 942;
// 128 x 512: 847 / 943 / 681 / 991; 96 x 640: 896 / -- / 766; 80 x 768: 951 / -- / 821; 64 x 1024: 996 / -- / 877; round 2, 168 x 384:
// 686 against 717 for 128 x 448 with the same cell-output kernel.
// Block size: the dependent cell-output grid (vicgpu_api.cu) never shares an SM with a resident step block in practice -- it runs
// on the SMs the step grid leaves idle -- so the step grid should be as small as the register file allows: 512-thread blocks put the
// 1,647 warps of the 10,000-cell domain on 103 SMs and leave 45 to the output grid (record time, winter week: 512 threads 553 us,
// 448 threads / 118 blocks 615 us, 384 / 138 blocks 609 us with the step slowed by co-resident output blocks; profiles/r02_summary.md).
#ifndef VICGPU_HRU_BLOCK_MAX
#define VICGPU_HRU_BLOCK_MAX 512
#endif
#ifndef VICGPU_STEP_MAXNREG
#define VICGPU_STEP_MAXNREG 128
#endif
#ifndef VICGPU_HRU_BLOCK
#define VICGPU_HRU_BLOCK 512
#endif

// one: the configuration's model step is a single sub-step (NF == 1)
void vicgpu_launch_hru_step_nn3(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0, int f_stride = 0, const int* block_w0 = nullptr, int nb_w0 = 0);
void vicgpu_launch_hru_step_nn10(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0, int f_stride = 0, const int* block_w0 = nullptr, int nb_w0 = 0);
void vicgpu_launch_hru_step_nn32(const vic::Opts* d_o, bool one, const vic::Tables& t, const double* frec, vic::Dmy d, int rec, vic::GlacAccum ga, int block, cudaStream_t s, unsigned long long* warp_ns = nullptr, long long sync_limit = 0, int nsm = 0, int f_stride = 0, const int* block_w0 = nullptr, int nb_w0 = 0);

int vicgpu_set_work_buffer_nn3(int* buf);
int vicgpu_set_work_buffer_nn10(int* buf);
int vicgpu_set_work_buffer_nn32(int* buf);
// measurement aid: record in, nframe live values in and out, record out (vicgpu_measure_phase_tax)
void vicgpu_launch_hru_pass_nn3(const vic::Opts* d_o, const vic::Tables& t, double* frame, int nframe, int block, cudaStream_t s);
void vicgpu_launch_hru_pass_nn10(const vic::Opts* d_o, const vic::Tables& t, double* frame, int nframe, int block, cudaStream_t s);
void vicgpu_launch_hru_pass_nn32(const vic::Opts* d_o, const vic::Tables& t, double* frame, int nframe, int block, cudaStream_t s);

#endif
