// Host-side internals shared by the translation units of libfiat_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "device_plan.cuh"

int fb_fail(int code, const std::string& msg);          // records the message for fiatb200_last_error()
extern std::atomic<long long> fb_launches;

#define FB_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fb_fail(FIATB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// Tuning overrides for experiments, read from the environment ONCE, the first time a launch
// asks for them; -1 = not set.  Nothing on the per-launch path calls getenv.
struct FbTuning {
    int mma_pt, mma_skip, mma_threads, tensor_bp, eval_bp, eval_generic, vals_tpc, vals_j, mma_wl, cells_reg, cells_threads;
};
const FbTuning& fb_tuning();

enum PlanKind { PLAN_SIMPLEX = 1, PLAN_TENSOR = 2, PLAN_LATTICE = 3 };

struct fiatb200_plan {
    int kind;
    int device;
    void* blob;             // one device allocation holding every table
    DevSimplex simplex;
    RecTab tab;             // host copy, passed to kernels by value
    SmallTab small_tab;     // generation-order coefficients for the register kernel
    DevTensor tensor;
    DevLattice lattice;
    int max_smem_optin;
    int num_sms;
    int max_segment;        // split-cell block streams: most blocks of one (row block, subcell) segment
    // staging for fiatb200_tabulate_host: two streams with one points/result buffer each, kept
    // across calls so that the end-to-end path issues no allocation or stream creation per call
    std::mutex host_mutex;
    cudaStream_t host_stream[2];
    double* host_pts[2];
    double* host_out[2];
    size_t host_pts_cap, host_out_cap;
};

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is a per-kernel maximum: it is only ever
// raised, under a lock, so that host threads launching the same kernel with different sizes cannot lower it
// under each other.
int fb_raise_smem_limit(const void* kernel, size_t bytes);
template <typename K>
int fb_set_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return FIATB200_OK;
    return fb_raise_smem_limit(reinterpret_cast<const void*>(kernel), bytes);
}

// launchers that live in their own translation units (compiled in parallel)
bool fb_small_applicable(const fiatb200_plan* plan);
int fb_dispatch_small(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                      double* out, long long ostride, const DevRowMap& M, cudaStream_t st);
bool fb_cells_applicable(const fiatb200_plan* plan);
int fb_dispatch_cells(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                      double* out, long long ostride, cudaStream_t st);
bool fb_vals_applicable(const fiatb200_plan* plan);
int fb_dispatch_vals(const fiatb200_plan* plan, const DevEntity& E, const double* pts, long long npts, long long ldp,
                     double* out, long long ostride, const DevRowMap& M, cudaStream_t st);
