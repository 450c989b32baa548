// Host-side launchers of the CAVI kernels (implemented in vl_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "vl_device.cuh"

// smallest supported number of m8 cluster tiles covering K clusters, or -1
int vl_supported_kt(int K);

cudaError_t vl_launch_contractions(int kt, int mode, const VlProblem* probs, const int2* htiles,
                                   int n_ht, const int2* ztiles, int n_zt, bool do_h, bool do_z,
                                   cudaStream_t stream);
cudaError_t vl_launch_scalar(const VlProblem* probs, int n_probs, int* done_count, cudaStream_t stream);
cudaError_t vl_launch_init(const VlProblem* probs, int n_probs, cudaStream_t stream);
cudaError_t vl_launch_mcount(const uint8_t* codes, double* mcount, int Npad, int Ls, cudaStream_t stream);
cudaError_t vl_launch_export_haplo(const double* h, double* out, int K, int L, int KS, cudaStream_t stream);
cudaError_t vl_launch_merge(const double* z, double* zm, const int* group_of, int N, int K, int G,
                            int KS, cudaStream_t stream);
