// Shared helpers for the tobac_flow_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tobac_flow_b200.h"

namespace tf {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return TF_ERR_CUDA;
    }
    return TF_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// host-side cvRound (round half to even)
inline int cv_round(double v) { return (int)nearbyint(v); }

constexpr int kMaxLevels = 8;

struct LevelPlan {
    int n;                 // number of levels processed (coarsest first)
    int h[kMaxLevels], w[kMaxLevels], k[kMaxLevels], ksize[kMaxLevels];
    double sigma[kMaxLevels];
};

LevelPlan make_level_plan(int H, int W, const tf_fb_params& p);

// streaming global loads/stores (data touched once per kernel: keep it out of L1)
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

}  // namespace tf
