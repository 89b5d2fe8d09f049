// K1: per-pair NaN-aware min/max and 8-bit quantisation.
// Replaces linear_norm + to_8bit(., 0, 1) (tobac_flow/utils/normalisation_utils.py:59-72, 10-33).
// HBM-bound: two streaming passes over the two fp32 frames (16 B/px) + 2 B/px written.
#include "tf_common.cuh"

namespace tf {

// order-preserving float <-> int key so atomicMin/atomicMax on ints give float min/max
__device__ __forceinline__ int f2key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* mm, int n_pairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pairs) {
        mm[2 * i] = f2key(INFINITY);
        mm[2 * i + 1] = f2key(-INFINITY);
    }
}

__device__ __forceinline__ void acc_minmax(float v, float& lo, float& hi) {
    // fminf/fmaxf ignore NaN operands == nanmin/nanmax
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
}

__global__ void __launch_bounds__(256) pair_minmax_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                          long long frame_stride, int n, int vec_ok,
                                                          int* __restrict__ mm) {
    const int p = blockIdx.y;
    const float* a = f0 + (long long)p * frame_stride;
    const float* b = f1 + (long long)p * frame_stride;
    float lo = INFINITY, hi = -INFINITY;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    if (vec_ok) {
        const int n4 = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int i = tid; i < n4; i += nthreads) {
            float4 u = __ldcs(a4 + i);
            float4 v = __ldcs(b4 + i);
            acc_minmax(u.x, lo, hi); acc_minmax(u.y, lo, hi); acc_minmax(u.z, lo, hi); acc_minmax(u.w, lo, hi);
            acc_minmax(v.x, lo, hi); acc_minmax(v.y, lo, hi); acc_minmax(v.z, lo, hi); acc_minmax(v.w, lo, hi);
        }
    } else {
        for (int i = tid; i < n; i += nthreads) {
            acc_minmax(a[i], lo, hi);
            acc_minmax(b[i], lo, hi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float slo[8], shi[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { slo[warp] = lo; shi[warp] = hi; }
    __syncthreads();
    if (warp == 0) {
        lo = lane < (blockDim.x >> 5) ? slo[lane] : INFINITY;
        hi = lane < (blockDim.x >> 5) ? shi[lane] : -INFINITY;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(mm + 2 * p, f2key(lo));
            atomicMax(mm + 2 * p + 1, f2key(hi));
        }
    }
}

// exact fp32 op order of the numpy expressions (no FMA contraction)
__device__ __forceinline__ float norm255(float x, float lo, float factor) {
    float a = __fmul_rn(__fsub_rn(x, lo), factor);
    // np.maximum(np.minimum(a, 1), 0) propagates NaN
    if (a == a) a = fmaxf(fminf(a, 1.0f), 0.0f);
    return __fmul_rn(a, 255.0f);
}

__device__ __forceinline__ void quantise_px(float x0, float x1, float lo, float factor, uint8_t& o0, uint8_t& o1) {
    float a0 = norm255(x0, lo, factor), a1 = norm255(x1, lo, factor);
    const bool k0 = isfinite(a0), k1 = isfinite(a1);
    if (!k0) a0 = 127.0f;
    if (!k1) a1 = 127.0f;
    if (!k0) a0 = a1;  // frame 0 takes frame 1's value (possibly the 127 fill)
    if (!k1) a1 = a0;  // frame 1 takes the (updated) frame 0 value
    o0 = (uint8_t)a0;  // astype(uint8): truncation
    o1 = (uint8_t)a1;
}

__global__ void __launch_bounds__(256) pair_quantise_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                            long long frame_stride, int n, int vec_ok,
                                                            const int* __restrict__ mm, uint8_t* __restrict__ q0,
                                                            uint8_t* __restrict__ q1) {
    const int p = blockIdx.y;
    const float* a = f0 + (long long)p * frame_stride;
    const float* b = f1 + (long long)p * frame_stride;
    uint8_t* oa = q0 + (long long)p * n;
    uint8_t* ob = q1 + (long long)p * n;
    const float lo = key2f(mm[2 * p]), hi = key2f(mm[2 * p + 1]);
    const float factor = (hi > lo) ? __fdiv_rn(1.0f, __fsub_rn(hi, lo)) : 0.0f;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    if (vec_ok) {
        const int n4 = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        uchar4* oa4 = reinterpret_cast<uchar4*>(oa);
        uchar4* ob4 = reinterpret_cast<uchar4*>(ob);
        for (int i = tid; i < n4; i += nthreads) {
            float4 u = __ldcs(a4 + i);
            float4 v = __ldcs(b4 + i);
            uchar4 r0, r1;
            quantise_px(u.x, v.x, lo, factor, r0.x, r1.x);
            quantise_px(u.y, v.y, lo, factor, r0.y, r1.y);
            quantise_px(u.z, v.z, lo, factor, r0.z, r1.z);
            quantise_px(u.w, v.w, lo, factor, r0.w, r1.w);
            oa4[i] = r0;
            ob4[i] = r1;
        }
    } else {
        for (int i = tid; i < n; i += nthreads) quantise_px(a[i], b[i], lo, factor, oa[i], ob[i]);
    }
}

// ---- float64 frames: the reference then normalises in float64 (numpy keeps the array dtype), which moves a few
// pixels across a truncation boundary relative to the float32 arithmetic above ----------------------------------------
__device__ __forceinline__ long long d2key(double d) {
    long long i = __double_as_longlong(d);
    return i >= 0 ? i : i ^ 0x7fffffffffffffffLL;
}
__device__ __forceinline__ double key2d(long long k) { return __longlong_as_double(k >= 0 ? k : k ^ 0x7fffffffffffffffLL); }

__global__ void minmax_init_f64_kernel(long long* mm, int n_pairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pairs) {
        mm[2 * i] = d2key((double)INFINITY);
        mm[2 * i + 1] = d2key(-(double)INFINITY);
    }
}

__global__ void __launch_bounds__(256) pair_minmax_f64_kernel(const double* __restrict__ f0, const double* __restrict__ f1,
                                                              long long frame_stride, int n, long long* __restrict__ mm) {
    const int p = blockIdx.y;
    const double* a = f0 + (long long)p * frame_stride;
    const double* b = f1 + (long long)p * frame_stride;
    double lo = INFINITY, hi = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double u = a[i], v = b[i];
        lo = fmin(fmin(lo, u), v);      // fmin / fmax ignore NaN operands == nanmin / nanmax
        hi = fmax(fmax(hi, u), v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * p, d2key(lo));
        atomicMax(mm + 2 * p + 1, d2key(hi));
    }
}

__global__ void __launch_bounds__(256) pair_quantise_f64_kernel(const double* __restrict__ f0, const double* __restrict__ f1,
                                                                long long frame_stride, int n,
                                                                const long long* __restrict__ mm, uint8_t* __restrict__ q0,
                                                                uint8_t* __restrict__ q1) {
    const int p = blockIdx.y;
    const double* a = f0 + (long long)p * frame_stride;
    const double* b = f1 + (long long)p * frame_stride;
    uint8_t* oa = q0 + (long long)p * n;
    uint8_t* ob = q1 + (long long)p * n;
    const double lo = key2d(mm[2 * p]), hi = key2d(mm[2 * p + 1]);
    const double factor = (hi > lo) ? __ddiv_rn(1.0, __dsub_rn(hi, lo)) : 0.0;
    auto norm255 = [&](double x) {
        double v = __dmul_rn(__dsub_rn(x, lo), factor);
        if (v == v) v = fmax(fmin(v, 1.0), 0.0);
        return __dmul_rn(v, 255.0);
    };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a0 = norm255(a[i]), a1 = norm255(b[i]);
        const bool k0 = isfinite(a0), k1 = isfinite(a1);
        if (!k0) a0 = 127.0;
        if (!k1) a1 = 127.0;
        if (!k0) a0 = a1;
        if (!k1) a1 = a0;
        oa[i] = (uint8_t)a0;
        ob[i] = (uint8_t)a1;
    }
}

}  // namespace tf

extern "C" int tf_pair_normalise_u8(const float* f0, const float* f1, long long frame_stride, uint8_t* q0, uint8_t* q1,
                                    int n_pairs, int H, int W, float* minmax_scratch, void* stream) {
    using namespace tf;
    if (n_pairs == 0) return TF_OK;
    if (!f0 || !f1 || !q0 || !q1 || !minmax_scratch || n_pairs < 0 || H <= 0 || W <= 0) {
        set_error("tf_pair_normalise_u8: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long n64 = (long long)H * W;
    if (n64 > 0x7fffffffLL) { set_error("tf_pair_normalise_u8: frame too large"); return TF_ERR_INVALID_ARGUMENT; }
    const int n = (int)n64;
    const int vec_ok = (n % 4 == 0) && (frame_stride % 4 == 0) && (((uintptr_t)f0 | (uintptr_t)f1) % 16 == 0) &&
                       (((uintptr_t)q0 | (uintptr_t)q1) % 4 == 0);
    int* mm = reinterpret_cast<int*>(minmax_scratch);
    LaunchTimer lt(KC_NORMALISE, 18.0 * n * n_pairs, s, 1 + 2 * cdiv(n_pairs, 65535));
    minmax_init_kernel<<<cdiv(n_pairs, 128), 128, 0, s>>>(mm, n_pairs);
    // enough blocks to fill 148 SMs x 8 resident blocks even for one pair; capped by the work available
    int bx = min(max(cdiv(n / 4, 256 * 4), 1), 1184);
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        int np = min(n_pairs - p0, 65535);
        dim3 grid(bx, np);
        pair_minmax_kernel<<<grid, 256, 0, s>>>(f0 + (long long)p0 * frame_stride, f1 + (long long)p0 * frame_stride,
                                                frame_stride, n, vec_ok, mm + 2 * p0);
        pair_quantise_kernel<<<grid, 256, 0, s>>>(f0 + (long long)p0 * frame_stride, f1 + (long long)p0 * frame_stride,
                                                  frame_stride, n, vec_ok, mm + 2 * p0, q0 + (long long)p0 * n,
                                                  q1 + (long long)p0 * n);
    }
    return check_launch("tf_pair_normalise_u8");
}

extern "C" int tf_pair_normalise_u8_f64(const double* f0, const double* f1, long long frame_stride, uint8_t* q0, uint8_t* q1,
                                        int n_pairs, int H, int W, double* minmax_scratch, void* stream) {
    using namespace tf;
    if (n_pairs == 0) return TF_OK;
    if (!f0 || !f1 || !q0 || !q1 || !minmax_scratch || n_pairs < 0 || H <= 0 || W <= 0) {
        set_error("tf_pair_normalise_u8_f64: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long n64 = (long long)H * W;
    if (n64 > 0x7fffffffLL) { set_error("tf_pair_normalise_u8_f64: frame too large"); return TF_ERR_INVALID_ARGUMENT; }
    const int n = (int)n64;
    long long* mm = reinterpret_cast<long long*>(minmax_scratch);
    LaunchTimer lt(KC_NORMALISE, 34.0 * n * n_pairs, s, 1 + 2 * cdiv(n_pairs, 65535));
    minmax_init_f64_kernel<<<cdiv(n_pairs, 128), 128, 0, s>>>(mm, n_pairs);
    const int bx = min(max(cdiv(n, 256 * 8), 1), 1184);
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int np = min(n_pairs - p0, 65535);
        dim3 grid(bx, np);
        pair_minmax_f64_kernel<<<grid, 256, 0, s>>>(f0 + (long long)p0 * frame_stride, f1 + (long long)p0 * frame_stride,
                                                    frame_stride, n, mm + 2 * p0);
        pair_quantise_f64_kernel<<<grid, 256, 0, s>>>(f0 + (long long)p0 * frame_stride, f1 + (long long)p0 * frame_stride,
                                                      frame_stride, n, mm + 2 * p0, q0 + (long long)p0 * n,
                                                      q1 + (long long)p0 * n);
    }
    return check_launch("tf_pair_normalise_u8_f64");
}
