// Host-side launchers of the CAVI kernels (implemented in vl_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "vl_device.cuh"

// Kernel classes by layout width: class 0 = up to 4 slot tiles, 1 = up to 8, 2 = up to 16.
#define VL_N_CLASSES 3
int vl_class_of_tiles(int kt);
int vl_class_capacity(int cls);

cudaError_t vl_launch_haplo(int cls, int mode, const VlProblem* probs, const int2* tiles, int n_tiles,
                            cudaStream_t stream);
cudaError_t vl_launch_cluster(int cls, int mode, const VlProblem* probs, const int2* tiles, int n_tiles,
                              cudaStream_t stream);
// plist entries: (problem, tile capacity of the class it is scheduled in); pstat[p] receives
// (still running, tiles needed next, valid slots, updates done)
cudaError_t vl_launch_scalar(const VlProblem* probs, const int2* plist, int n, int* done_count, int4* pstat,
                             cudaStream_t stream);
cudaError_t vl_launch_init(const VlProblem* probs, int n_probs, int4* pstat, cudaStream_t stream);
cudaError_t vl_launch_prepare(const uint8_t* codes, const double* lt, const double* lo, const uint8_t* maj,
                              double* dm, double* mcount, double* ltsum, int N, int L, int Npad, int Ls,
                              int mode, cudaStream_t stream);
cudaError_t vl_launch_export_haplo(const VlProblem* prob, double* out, size_t total, cudaStream_t stream);
cudaError_t vl_launch_export_cluster(const VlProblem* prob, double* out, size_t total, cudaStream_t stream);
cudaError_t vl_launch_merge(const VlProblem* prob, double* zm, const int* group_of, int G, int KSg,
                            size_t total, cudaStream_t stream);
