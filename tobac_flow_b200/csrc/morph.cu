// K12: the per-frame filters detect_growth_markers applies around the Flow operators (tobac_flow/detection.py:64-125),
// with scipy.ndimage's exact semantics so the marker masks stay bit-identical:
//   * gaussian_filter(field, (0, sigma, sigma))      -> gaussian_y_kernel + gaussian_x_kernel (fp64 accumulation in
//     scipy's correlate1d order: centre first, then symmetric pairs from the outside in; 'reflect' borders; the
//     intermediate is rounded to the array dtype between the two passes exactly as scipy does);
//   * the second-difference curvature test           -> curvature_mask_kernel;
//   * binary_opening with the 2-D cross              -> binary_opening_cross_kernel (erosion then dilation, border 0);
//   * grey_opening with the 2-D cross footprint      -> grey_erode/dilate_cross_kernel ('reflect' borders; scipy's
//     min/max scan: the running value starts at the first footprint element, later elements replace it only when the
//     comparison is true, so a NaN propagates only from the first element);
//   * per-frame division by the time step            -> scale_frames_kernel.
#include "tf_common.cuh"

namespace tf {

constexpr int kMaxGaussRadius = 32;
struct GaussTaps {
    int radius;
    double w[kMaxGaussRadius + 1];   // w[0] = centre, w[k] = weight at distance k
};

// scipy 'reflect' (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_sym(int i, int n) {
    if (n == 1) return 0;
    const int n2 = 2 * n;
    i %= n2;
    if (i < 0) i += n2;
    return i < n ? i : n2 - 1 - i;
}

template <typename T>
__global__ void __launch_bounds__(256) gaussian_y_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W,
                                                         GaussTaps g) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    const T* col = in + base + x;
    double acc = __dmul_rn((double)col[(long long)y * W], g.w[0]);   // no FMA contraction: scipy's C loop has none
    const bool interior = y - g.radius >= 0 && y + g.radius < H;
    for (int k = g.radius; k >= 1; --k) {
        const int ya = interior ? y - k : reflect_sym(y - k, H), yb = interior ? y + k : reflect_sym(y + k, H);
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)col[(long long)ya * W], (double)col[(long long)yb * W]), g.w[k]));
    }
    out[base + (long long)y * W + x] = (T)acc;
}

template <typename T>
__global__ void __launch_bounds__(256) gaussian_x_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W,
                                                         GaussTaps g) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W + (long long)y * W;
    const T* row = in + base;
    double acc = __dmul_rn((double)row[x], g.w[0]);
    const bool interior = x - g.radius >= 0 && x + g.radius < W;
    for (int k = g.radius; k >= 1; --k) {
        const int xa = interior ? x - k : reflect_sym(x - k, W), xb = interior ? x + k : reflect_sym(x + k, W);
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)row[xa], (double)row[xb]), g.w[k]));
    }
    out[base + x] = (T)acc;
}

// np.diff(s, n=2) along x and y (in the array dtype), zero on the border rows / columns, both beyond the threshold
template <typename T>
__global__ void __launch_bounds__(256) curvature_mask_kernel(const T* __restrict__ s, uint8_t* __restrict__ out, int H, int W,
                                                             double threshold, int positive) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    const T* p = s + base + (long long)y * W + x;
    double xd = 0.0, yd = 0.0;
    const T c = p[0];
    if (x > 0 && x < W - 1) xd = (double)((T)(p[1] - c) - (T)(c - p[-1]));
    if (y > 0 && y < H - 1) yd = (double)((T)(p[W] - c) - (T)(c - p[-W]));
    const bool m = positive ? (xd > threshold && yd > threshold) : (xd < -threshold && yd < -threshold);
    out[base + (long long)y * W + x] = m ? 1 : 0;
}

// binary_opening(structure = cross): erosion (outside = 0) followed by dilation (outside = 0)
__global__ void __launch_bounds__(256) binary_opening_cross_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                                   int H, int W) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    const uint8_t* m = in + base;
    auto at = [&](int yy, int xx) -> bool {
        return (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W && m[(long long)yy * W + xx] != 0;
    };
    auto eroded = [&](int yy, int xx) -> bool {
        if (!((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W)) return false;
        return at(yy, xx) && at(yy - 1, xx) && at(yy + 1, xx) && at(yy, xx - 1) && at(yy, xx + 1);
    };
    const bool o = eroded(y, x) || eroded(y - 1, x) || eroded(y + 1, x) || eroded(y, x - 1) || eroded(y, x + 1);
    out[base + (long long)y * W + x] = o ? 1 : 0;
}

// minimum_filter / maximum_filter with the cross footprint, scipy scan order (y-1, x-1, centre, x+1, y+1)
template <typename T, bool IS_MIN>
__global__ void __launch_bounds__(256) grey_cross_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    const T* a = in + base;
    const int ym = reflect_sym(y - 1, H), yp = reflect_sym(y + 1, H), xm = reflect_sym(x - 1, W), xp = reflect_sym(x + 1, W);
    T v = a[(long long)ym * W + x];
    const T t1 = a[(long long)y * W + xm], t2 = a[(long long)y * W + x], t3 = a[(long long)y * W + xp],
            t4 = a[(long long)yp * W + x];
    if (IS_MIN) {
        if (t1 < v) v = t1;
        if (t2 < v) v = t2;
        if (t3 < v) v = t3;
        if (t4 < v) v = t4;
    } else {
        if (t1 > v) v = t1;
        if (t2 > v) v = t2;
        if (t3 > v) v = t3;
        if (t4 > v) v = t4;
    }
    out[base + (long long)y * W + x] = v;
}

// out[t] = (double)in[t] / dt[t]   (Flow.diff(wvd) / get_time_diff_from_coord(wvd.t)[:, None, None], detection.py:99-101)
__global__ void __launch_bounds__(256) scale_frames_kernel(const float* __restrict__ in, const double* __restrict__ dt,
                                                           double* __restrict__ out, long long hw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const long long o = (long long)blockIdx.y * hw + i;
    out[o] = (double)in[o] / dt[blockIdx.y];
}

// out = a * (mask != 0)  in a's dtype   (grey_opening(...) * get_curvature_filter(...), detection.py:105-108)
template <typename T>
__global__ void __launch_bounds__(256) mask_multiply_kernel(const T* __restrict__ a, const uint8_t* __restrict__ m,
                                                            T* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = a[i] * (T)(m[i] ? 1 : 0);   // NaN * 0 = NaN, as numpy
}

// out = (a >= thr) as uint8   (comparisons with NaN are false, as numpy)
template <typename T>
__global__ void __launch_bounds__(256) threshold_ge_kernel(const T* __restrict__ a, double thr, uint8_t* __restrict__ out,
                                                           long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = ((double)a[i] >= thr) ? 1 : 0;
}

static int frames_ok(const char* who, const void* a, const void* b, int T, int H, int W) {
    if (!a || !b || T < 0 || H <= 0 || W <= 0) { set_error("%s: invalid argument", who); return TF_ERR_INVALID_ARGUMENT; }
    if (T > 65535) { set_error("%s: more than 65535 frames per call", who); return TF_ERR_UNSUPPORTED; }
    if ((long long)H * W > 0x7fffffffLL) { set_error("%s: frame too large", who); return TF_ERR_INVALID_ARGUMENT; }
    return TF_OK;
}

}  // namespace tf

using namespace tf;

extern "C" int tf_gaussian_filter_yx(const void* in, void* tmp, void* out, int dtype, int T, int H, int W,
                                     const double* weights, int radius, void* stream) {
    if (T == 0) return TF_OK;
    int rc = frames_ok("tf_gaussian_filter_yx", in, out, T, H, W);
    if (rc != TF_OK) return rc;
    if (!tmp || !weights || radius < 0) { set_error("tf_gaussian_filter_yx: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    if (dtype != TF_F32 && dtype != TF_F64) { set_error("tf_gaussian_filter_yx: float32 / float64 only"); return TF_ERR_UNSUPPORTED; }
    if (radius > kMaxGaussRadius) { set_error("tf_gaussian_filter_yx: radius %d too large", radius); return TF_ERR_UNSUPPORTED; }
    GaussTaps g;
    g.radius = radius;
    for (int k = 0; k <= radius; ++k) g.w[k] = weights[radius + k];   // symmetric kernel: centre and one side
    cudaStream_t s = (cudaStream_t)stream;
    const double es = dtype == TF_F32 ? 4.0 : 8.0;
    LaunchTimer lt(KC_MORPH, 4.0 * es * H * W * T, s, 2);
    dim3 blk(32, 8), grid(cdiv(W, 32), cdiv(H, 8), T);
    if (dtype == TF_F32) {
        gaussian_y_kernel<float><<<grid, blk, 0, s>>>((const float*)in, (float*)tmp, H, W, g);
        gaussian_x_kernel<float><<<grid, blk, 0, s>>>((const float*)tmp, (float*)out, H, W, g);
    } else {
        gaussian_y_kernel<double><<<grid, blk, 0, s>>>((const double*)in, (double*)tmp, H, W, g);
        gaussian_x_kernel<double><<<grid, blk, 0, s>>>((const double*)tmp, (double*)out, H, W, g);
    }
    return check_launch("tf_gaussian_filter_yx");
}

extern "C" int tf_curvature_mask(const void* smoothed, uint8_t* out, int dtype, int T, int H, int W, double threshold,
                                 int positive, void* stream) {
    if (T == 0) return TF_OK;
    int rc = frames_ok("tf_curvature_mask", smoothed, out, T, H, W);
    if (rc != TF_OK) return rc;
    if (dtype != TF_F32 && dtype != TF_F64) { set_error("tf_curvature_mask: float32 / float64 only"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_MORPH, ((dtype == TF_F32 ? 4.0 : 8.0) + 1.0) * H * W * T, s, 1);
    dim3 blk(32, 8), grid(cdiv(W, 32), cdiv(H, 8), T);
    if (dtype == TF_F32) curvature_mask_kernel<float><<<grid, blk, 0, s>>>((const float*)smoothed, out, H, W, threshold, positive);
    else curvature_mask_kernel<double><<<grid, blk, 0, s>>>((const double*)smoothed, out, H, W, threshold, positive);
    return check_launch("tf_curvature_mask");
}

extern "C" int tf_binary_opening_cross(const uint8_t* in, uint8_t* out, int T, int H, int W, void* stream) {
    if (T == 0) return TF_OK;
    int rc = frames_ok("tf_binary_opening_cross", in, out, T, H, W);
    if (rc != TF_OK) return rc;
    if (in == out) { set_error("tf_binary_opening_cross: output must not alias the input"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_MORPH, 2.0 * H * W * T, s, 1);
    dim3 blk(32, 8), grid(cdiv(W, 32), cdiv(H, 8), T);
    binary_opening_cross_kernel<<<grid, blk, 0, s>>>(in, out, H, W);
    return check_launch("tf_binary_opening_cross");
}

extern "C" int tf_grey_opening_cross(const void* in, void* tmp, void* out, int dtype, int T, int H, int W, void* stream) {
    if (T == 0) return TF_OK;
    int rc = frames_ok("tf_grey_opening_cross", in, out, T, H, W);
    if (rc != TF_OK) return rc;
    if (!tmp || tmp == in || tmp == out) { set_error("tf_grey_opening_cross: tmp must be a distinct buffer"); return TF_ERR_INVALID_ARGUMENT; }
    if (dtype != TF_F32 && dtype != TF_F64) { set_error("tf_grey_opening_cross: float32 / float64 only"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_MORPH, 4.0 * (dtype == TF_F32 ? 4.0 : 8.0) * H * W * T, s, 2);
    dim3 blk(32, 8), grid(cdiv(W, 32), cdiv(H, 8), T);
    if (dtype == TF_F32) {
        grey_cross_kernel<float, true><<<grid, blk, 0, s>>>((const float*)in, (float*)tmp, H, W);
        grey_cross_kernel<float, false><<<grid, blk, 0, s>>>((const float*)tmp, (float*)out, H, W);
    } else {
        grey_cross_kernel<double, true><<<grid, blk, 0, s>>>((const double*)in, (double*)tmp, H, W);
        grey_cross_kernel<double, false><<<grid, blk, 0, s>>>((const double*)tmp, (double*)out, H, W);
    }
    return check_launch("tf_grey_opening_cross");
}

extern "C" int tf_scale_frames(const float* in, const double* dt, double* out, int T, int H, int W, void* stream) {
    if (T == 0) return TF_OK;
    int rc = frames_ok("tf_scale_frames", in, out, T, H, W);
    if (rc != TF_OK) return rc;
    if (!dt) { set_error("tf_scale_frames: dt is NULL"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    const long long hw = (long long)H * W;
    LaunchTimer lt(KC_MORPH, 12.0 * hw * T, s, 1);
    dim3 grid((unsigned)((hw + 255) / 256), T);
    scale_frames_kernel<<<grid, 256, 0, s>>>(in, dt, out, hw);
    return check_launch("tf_scale_frames");
}

extern "C" int tf_mask_multiply(const void* a, const uint8_t* mask, void* out, int dtype, long long n, void* stream) {
    if (n == 0) return TF_OK;
    if (!a || !mask || !out || n < 0) { set_error("tf_mask_multiply: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    if (dtype != TF_F32 && dtype != TF_F64) { set_error("tf_mask_multiply: float32 / float64 only"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_MORPH, (2.0 * (dtype == TF_F32 ? 4.0 : 8.0) + 1.0) * n, s, 1);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == TF_F32) mask_multiply_kernel<float><<<blocks, 256, 0, s>>>((const float*)a, mask, (float*)out, n);
    else mask_multiply_kernel<double><<<blocks, 256, 0, s>>>((const double*)a, mask, (double*)out, n);
    return check_launch("tf_mask_multiply");
}

extern "C" int tf_threshold_ge(const void* a, double threshold, uint8_t* out, int dtype, long long n, void* stream) {
    if (n == 0) return TF_OK;
    if (!a || !out || n < 0) { set_error("tf_threshold_ge: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    if (dtype != TF_F32 && dtype != TF_F64) { set_error("tf_threshold_ge: float32 / float64 only"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_MORPH, ((dtype == TF_F32 ? 4.0 : 8.0) + 1.0) * n, s, 1);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (dtype == TF_F32) threshold_ge_kernel<float><<<blocks, 256, 0, s>>>((const float*)a, threshold, out, n);
    else threshold_ge_kernel<double><<<blocks, 256, 0, s>>>((const double*)a, threshold, out, n);
    return check_launch("tf_threshold_ge");
}
