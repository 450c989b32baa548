// Host-side launchers of the CAVI kernels (implemented in vl_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "vl_device.cuh"

// Kernel classes by layout width: class c handles up to 2 << c slot tiles (2, 4, 8, 16); the
// narrowest one is compiled for 8 CTAs per SM.
#define VL_N_CLASSES 4
int vl_class_of_tiles(int kt);
int vl_class_capacity(int cls);

cudaError_t vl_launch_haplo(int cls, int mode, int split, const VlProblem* probs, const int2* tiles, int n_tiles,
                            cudaStream_t stream);
// pstat[p] receives (still running, tiles needed next, valid slots, updates done) from the scalar
// update that the last read tile of restart p performs
cudaError_t vl_launch_cluster(int cls, int mode, const VlProblem* probs, const int2* tiles, int n_tiles,
                              int* done_count, int4* pstat, cudaStream_t stream);
// n_updates updates of the restarts rlist[0..n_r) of one kernel class in one launch of n_blocks one-tile
// workers (vl_sched_kernel; n_blocks = tiles the chunk can have): chunk_base[p] = (updates restart p had
// completed when the chunk was scheduled, its haplotype tiles x read parts, its read tiles, 0), ctl[p] zeroed and *n_live = n_r set by the caller; deep = large
// shared memory per block (few tiles in flight)
cudaError_t vl_launch_sched(int cls, int mode, int split, const VlProblem* probs, const int* rlist, int n_r, int n_updates,
                            const int4* chunk_base, VlCtl* ctl, int* n_live, int* done_count, int4* pstat, long long n_blocks,
                            int deep, int persistent, cudaStream_t stream);
cudaError_t vl_launch_init(const VlProblem* probs, int n_probs, int4* pstat, cudaStream_t stream);
cudaError_t vl_launch_prepare(const uint8_t* codes, const double* lt, const double* lo, const int* col_lb,
                              double* dm, double* mcount, double* ltsum, int N, int L, int Npad, int NCs,
                              int mode, cudaStream_t stream);
cudaError_t vl_launch_export_haplo(const VlProblem* prob, double* out, size_t total, cudaStream_t stream);
cudaError_t vl_launch_export_cluster(const VlProblem* prob, double* out, size_t total, cudaStream_t stream);
cudaError_t vl_launch_merge(const VlProblem* prob, double* zm, double* wzm, const int* group_of, int G, int KSg,
                            size_t total, cudaStream_t stream);
cudaError_t vl_launch_scale_rows(const double* z, const double* w, double* wz, int Npad, int KS, cudaStream_t stream);
cudaError_t vl_launch_init_rows(const double* raw, const double* w, double* z, double* wz, int N, int K, int Npad, int KS,
                                cudaStream_t stream);
cudaError_t vl_launch_exp_test(const double* x, double* out, int n, cudaStream_t stream);
// out: n_blocks * 256 doubles; every warp issues 8 * iters DMMAs
cudaError_t vl_launch_dmma_peak(double* out, int n_blocks, int iters, cudaStream_t stream);
// per-phase cycle counters of restart 0 in the general-path kernels (see vl_kernels.cu); 2 unless -DVL_RS_PROFILE
// records a message for vl_last_error and returns 1 (vl_api.cu)
int vl_set_error(const char* msg);
// id of a haplotype tile in the tile lists: column tile, which part of the reads, of how many parts
inline int vl_htile_id(int tile, int split, int n_split) { return n_split > 1 ? (tile | (split << 20) | (n_split << 24)) : tile; }
int vl_debug_general_cycles(unsigned long long* out20, int reset);
cudaError_t vl_sched_occupancy(int cls, int mode, int split, int deep, int persistent, int* ctas_per_sm, int* smem_bytes, int* regs);
