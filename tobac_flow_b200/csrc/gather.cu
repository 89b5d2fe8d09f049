// K8: semi-Lagrangian 3x3x3 tap gather with fused per-step reducers; K9: smooth_flow_step; K7: finalise.
//
// Replaces tobac_flow/convolve.py (warp_flow -> cv2.remap with BORDER_CONSTANT, convolve_same_step, convolve_step,
// convolve) and the reducers of Flow.diff (flow.py:182-186), sobel.py and detection.py.  cv2.remap semantics are
// reproduced exactly (see oracle/remap_np.py): 1/32-px coordinate quantisation with round-half-even, fp32 table
// weights wy*wx, left-to-right accumulation without FMA contraction, per-tap constant border (NaN poisoning through
// zero weights), short saturation of integer coordinates.
//
// One thread per output pixel; the tap stack lives in registers only.  HBM traffic per pixel-step: 3 operand frames
// (neighbour re-reads hit L1/L2), 2 flow fields (16 B) and the result.
#include <stdlib.h>

#include <mutex>
#include <type_traits>

#include "tf_common.cuh"

namespace tf {

enum RedClass { RC_NONE = 0, RC_DIFF = 1, RC_STAT = 2, RC_SOBEL = 3 };

struct GatherArgs {
    const void* cur0;
    const float2* fflow0;
    const float2* bflow0;
    void* out;
    long long out_tap_stride;
    int n_frames, has_prev, has_next;
    int H, W;
    unsigned structure;  // bit k = tap k present, k = t*9 + y*3 + x
    int reducer;         // tf_reducer
    double fill;
};

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

// cvRound on a float coordinate with x86 semantics for NaN/overflow (-> INT_MIN), then saturate_cast<short>
__device__ __forceinline__ int cv_round_dev(float v) {
    if (!(fabsf(v) < 2.0e9f)) return INT_MIN;
    return __float2int_rn(v);
}
__device__ __forceinline__ int sat_short(int v) { return min(max(v, -32768), 32767); }

// rint(32 p) without the conversion unit: 32 p is exact in fp32, so the single rounding of fma(p, 32, 1.5 * 2^23) IS
// cvRound's round-half-even, and the integer sits in the low mantissa bits.  Exact for |32 p| < 2^22; outside that range
// (and for NaN / infinities) the returned value is negative or >= 2^22, i.e. it fails every "footprint inside the image"
// test below (images are at most 32767 wide: cv2.remap's short coordinates), and the caller takes the general path.
constexpr float kQMagic = 12582912.f;
constexpr int kQMagicBits = 0x4B400000;
__device__ __forceinline__ int quantise_fast(float p) { return __float_as_int(fmaf(p, 32.f, kQMagic)) - kQMagicBits; }

// a source image that may be a real frame or the virtual all-`fill` frame of convolve.py:307-314
template <typename T>
struct FrameSrc {
    const T* p;   // nullptr -> constant frame
    T fill;
    int H, W;
    __device__ __forceinline__ T at(int y, int x) const { return p ? p[y * W + x] : fill; }
};

// strided view (component c of an interleaved (H, W, 2) flow field) for smooth_flow_step
struct FlowCompSrc {
    const float* p;
    float fill;
    int H, W;
    __device__ __forceinline__ float at(int y, int x) const { return p[(y * W + x) * 2]; }
};

__device__ __forceinline__ void cubic_coeffs(int fi, float c[4]) {
    // interpolateCubic, A = -0.75, fp32, no contraction
    const float A = -0.75f;
    const float x = __fmul_rn((float)fi, 1.0f / 32.0f);
    const float xp1 = __fadd_rn(x, 1.f);
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), __fmul_rn(5.f, A)), xp1), __fmul_rn(8.f, A)), xp1),
                     __fmul_rn(4.f, A));
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f)), x), x), 1.f);
    const float omx = __fsub_rn(1.f, x);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), omx), __fadd_rn(A, 3.f)), omx), omx), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// cv2.INTER_LANCZOS4: the 32 x 8 fp32 coefficient table of interpolateLanczos4 (imgwarp.cpp), built on the host exactly as
// OpenCV builds it (sines / cosines in double from the same libm, normalisation in fp32) and kept in constant memory
__constant__ float c_lanczos4[32 * 8];

// cv2.remap at one position.  T: element type, Src: accessor with at(y, x), H, W, fill
template <int INTERP, typename T, typename Src>
__device__ __forceinline__ T remap_at(const Src& s, float px, float py) {
    const int H = s.H, W = s.W;
    if constexpr (INTERP == TF_NEAREST) {
        const int ix = sat_short(cv_round_dev(px)), iy = sat_short(cv_round_dev(py));
        if ((unsigned)ix < (unsigned)W && (unsigned)iy < (unsigned)H) return s.at(iy, ix);
        return s.fill;
    } else {
        if constexpr (INTERP == TF_LINEAR) {
            // the 2 x 2 footprint inside the image (almost every sample): no float -> int conversions, one range test per axis
            const int qx = quantise_fast(px), qy = quantise_fast(py);
            if ((unsigned)qx < (unsigned)(min(max(W - 1, 0), 32767) << 5) && (unsigned)qy < (unsigned)(min(max(H - 1, 0), 32767) << 5)) {
                const int ix = qx >> 5, iy = qy >> 5;
                const float fx = (float)(qx & 31) * (1.0f / 32.0f), fy = (float)(qy & 31) * (1.0f / 32.0f);
                const float wx0 = 1.f - fx, wy0 = 1.f - fy;  // exact
                const T w00 = (T)(wy0 * wx0), w01 = (T)(wy0 * fx), w10 = (T)(fy * wx0), w11 = (T)(fy * fx);  // exact
                const T v00 = s.at(iy, ix), v01 = s.at(iy, ix + 1), v10 = s.at(iy + 1, ix), v11 = s.at(iy + 1, ix + 1);
                return add_rn(add_rn(add_rn(mul_rn(v00, w00), mul_rn(v01, w01)), mul_rn(v10, w10)), mul_rn(v11, w11));
            }
        }
        const int sx = cv_round_dev(__fmul_rn(px, 32.f)), sy = cv_round_dev(__fmul_rn(py, 32.f));
        const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5);
        const int fxi = sx & 31, fyi = sy & 31;
        if constexpr (INTERP == TF_LINEAR) {
            const float fx = (float)fxi * (1.0f / 32.0f), fy = (float)fyi * (1.0f / 32.0f);
            const float wx0 = 1.f - fx, wy0 = 1.f - fy;  // exact
            const T w00 = (T)(wy0 * wx0), w01 = (T)(wy0 * fx), w10 = (T)(fy * wx0), w11 = (T)(fy * fx);  // exact
            // (footprints inside the image were taken above; this is the border form, valid for any position)
            if (ix >= W || ix + 1 < 0 || iy >= H || iy + 1 < 0) return s.fill;
            const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
            const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
            const T v00 = (y0 && x0) ? s.at(iy, ix) : s.fill;
            const T v01 = (y0 && x1) ? s.at(iy, ix + 1) : s.fill;
            const T v10 = (y1 && x0) ? s.at(iy + 1, ix) : s.fill;
            const T v11 = (y1 && x1) ? s.at(iy + 1, ix + 1) : s.fill;
            return add_rn(add_rn(add_rn(mul_rn(v00, w00), mul_rn(v01, w01)), mul_rn(v10, w10)), mul_rn(v11, w11));
        } else if constexpr (INTERP == TF_LANCZOS4) {
            // remapLanczos4: 8 x 8 taps from (ix - 3, iy - 3); weights wy[k1] * wx[k2] (fp32 table product); inside: the
            // eight-term row expression left to right, rows added in order; near the border: cv + sum (S - cv) * w
            const float* cx = c_lanczos4 + 8 * fxi;
            const float* cy = c_lanczos4 + 8 * fyi;
            const int x0 = ix - 3, y0 = iy - 3;
            if ((unsigned)x0 < (unsigned)max(W - 7, 0) && (unsigned)y0 < (unsigned)max(H - 7, 0)) {
                T sum = (T)0;
#pragma unroll 1
                for (int k1 = 0; k1 < 8; ++k1) {
                    T row = mul_rn(s.at(y0 + k1, x0), (T)__fmul_rn(cy[k1], cx[0]));
#pragma unroll
                    for (int k2 = 1; k2 < 8; ++k2)
                        row = add_rn(row, mul_rn(s.at(y0 + k1, x0 + k2), (T)__fmul_rn(cy[k1], cx[k2])));
                    sum = add_rn(sum, row);
                }
                return sum;
            }
            if (x0 >= W || x0 + 8 <= 0 || y0 >= H || y0 + 8 <= 0) return s.fill;
            const T cv = s.fill;
            T sum = cv;
#pragma unroll 1
            for (int k1 = 0; k1 < 8; ++k1) {
                const int yy = y0 + k1;
                if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll 1
                for (int k2 = 0; k2 < 8; ++k2) {
                    const int xx = x0 + k2;
                    if ((unsigned)xx < (unsigned)W)
                        sum = add_rn(sum, mul_rn(sub_rn(s.at(yy, xx), cv), (T)__fmul_rn(cy[k1], cx[k2])));
                }
            }
            return sum;
        } else {  // cubic
            float cx[4], cy[4];
            cubic_coeffs(fxi, cx);
            cubic_coeffs(fyi, cy);
            const int x0 = ix - 1, y0 = iy - 1;
            if ((unsigned)x0 < (unsigned)max(W - 3, 0) && (unsigned)y0 < (unsigned)max(H - 3, 0)) {
                T sum = (T)0;
#pragma unroll
                for (int k1 = 0; k1 < 4; ++k1) {
                    T row = mul_rn(s.at(y0 + k1, x0), (T)__fmul_rn(cy[k1], cx[0]));
#pragma unroll
                    for (int k2 = 1; k2 < 4; ++k2)
                        row = add_rn(row, mul_rn(s.at(y0 + k1, x0 + k2), (T)__fmul_rn(cy[k1], cx[k2])));
                    sum = k1 == 0 ? row : add_rn(sum, row);
                }
                return sum;
            }
            if (x0 >= W || x0 + 4 <= 0 || y0 >= H || y0 + 4 <= 0) return s.fill;
            const T cv = s.fill;
            T sum = cv;  // cv * ONE
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
                const int yy = y0 + k1;
                if ((unsigned)yy >= (unsigned)H) continue;
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const int xx = x0 + k2;
                    if ((unsigned)xx < (unsigned)W)
                        sum = add_rn(sum, mul_rn(sub_rn(s.at(yy, xx), cv), (T)__fmul_rn(cy[k1], cx[k2])));
                }
            }
            return sum;
        }
    }
}

template <typename T> __device__ __forceinline__ bool is_nan(T v) { return v != v; }
template <> __device__ __forceinline__ bool is_nan<int>(int) { return false; }
template <typename T> __device__ __forceinline__ bool is_fin(T v) { return isfinite(v); }
template <> __device__ __forceinline__ bool is_fin<int>(int) { return true; }

template <typename S, typename D> __device__ __forceinline__ D cast_to(S v) { return (D)v; }

template <typename T> __device__ __forceinline__ T nan_of() { return (T)NAN; }
template <> __device__ __forceinline__ int nan_of<int>() { return 0; }

template <typename T> __device__ __forceinline__ T fill_cast(double f) { return (T)f; }
template <> __device__ __forceinline__ int fill_cast<int>(double f) {
    if (!(f == f)) return 0;
    return (int)f;
}

// ------------------------------------------------------------------------------------------------------------------
// reducers: fed taps in stack order (t-1 slab, t slab, t+1 slab; row-major inside), value already in stack dtype
// ------------------------------------------------------------------------------------------------------------------
template <typename ST>
struct RedNone {
    ST* out; long long tap_stride; long long pix;
    __device__ __forceinline__ void add(int n, int, ST v) { out[n * tap_stride + pix] = v; }
};

template <typename ST>
struct RedDiff {
    ST x[3];
    __device__ __forceinline__ void add(int n, int, ST v) { if (n < 3) x[n] = v; }
    __device__ __forceinline__ ST finish() const {
        const ST a = x[2] - x[1], b = x[1] - x[0];
        const ST s = (is_nan(a) ? (ST)0 : a) + (is_nan(b) ? (ST)0 : b);
        int cnt = (is_fin(x[2]) ? 1 : 0) + (is_fin(x[0]) ? 1 : 0);
        cnt = max(cnt, 1);
        // numpy divides through fp64 and rounds back; dividing by 1 or 2 is exact in any precision
        return cnt == 2 ? s * (ST)0.5 : s;
    }
};

template <typename ST>
struct RedStat {  // nanmean / nanmax / nanmin / any
    int mode; ST acc; int cnt; bool any;
    __device__ __forceinline__ void init(int m) { mode = m; acc = (ST)0; cnt = 0; any = false; }
    __device__ __forceinline__ void add(int, int, ST v) {
        if (v != (ST)0) any = true;  // NaN != 0 is true, like np.any on floats
        if (is_nan(v)) return;
        if (mode == TF_RED_NANMEAN) acc = acc + v;
        else if (mode == TF_RED_NANMAX) acc = cnt ? (v > acc ? v : acc) : v;
        else if (mode == TF_RED_NANMIN) acc = cnt ? (v < acc ? v : acc) : v;
        ++cnt;
    }
    __device__ __forceinline__ ST finish() const {
        if (mode == TF_RED_ANY) return any ? (ST)1 : (ST)0;
        if (mode == TF_RED_NANMEAN) return (ST)((double)acc / (double)cnt);  // 0/0 -> NaN like numpy
        return cnt ? acc : nan_of<ST>();
    }
};

// Sobel magnitude over the 27 taps (sobel.py:7-86): taps are kept in registers in the stack dtype and reduced at the end.
// KT: type the taps are kept in = the narrower of operand and stack dtype (the cast to ST is then exact and is redone on use)
template <typename ST, typename KT>
struct RedSobel {
    int dir; ST centre; KT v[27];
    __device__ __forceinline__ void init(int d, ST c) { dir = d; centre = c; }
    __device__ __forceinline__ void add(int, int k, KT val) { v[k] = val; }
    // general form: d_k = v_k - centre (optionally clipped), NaN terms skipped, fp64 sums of d_k * S_k
    __device__ __forceinline__ ST finish_general() const {
        double gx = 0.0, gy = 0.0, gt = 0.0;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            ST d = (ST)v[k] - centre;
            if (dir == TF_RED_SOBEL_UPHILL) d = is_nan(d) ? (ST)0 : (d > (ST)0 ? d : (ST)0);        // np.fmax(d, 0)
            else if (dir == TF_RED_SOBEL_DOWNHILL) d = is_nan(d) ? (ST)0 : (d < (ST)0 ? d : (ST)0); // np.fmin(d, 0)
            if (is_nan(d)) continue;
            const int t = k / 9, y = (k / 3) % 3, x = k % 3;
            const int wt = 2 - (t - 1) * (t - 1), wy = 2 - (y - 1) * (y - 1), wx = 2 - (x - 1) * (x - 1);  // (1, 2, 1)
            const double dd = (double)d;
            if (wt * wy * (x - 1) != 0) gx += dd * (double)(wt * wy * (x - 1));
            if (wx * wt * (y - 1) != 0) gy += dd * (double)(wx * wt * (y - 1));
            if (wy * wx * (t - 1) != 0) gt += dd * (double)(wy * wx * (t - 1));
        }
        return (ST)sqrt(gx * gx + gy * gy + gt * gt);
    }
    // plain direction, fp64 stack, no NaN tap: the weights of each gradient sum to zero, so the centre drops out and
    // the (1,2,1) x (1,2,1) x (-1,0,1) kernels are applied separably to the raw taps.  Every first-level sum /
    // difference of fp32-valued taps is exact in fp64, so this is at least as accurate as the general form.
    __device__ __forceinline__ double finish_separable() const { return separable(v); }
    __device__ __forceinline__ static double separable(const KT* v) {
        double gxs[3], gys[3], gts[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            double rd[3], rs[3];
#pragma unroll
            for (int y = 0; y < 3; ++y) {
                const double a = (double)v[t * 9 + y * 3], b = (double)v[t * 9 + y * 3 + 1], c = (double)v[t * 9 + y * 3 + 2];
                rd[y] = c - a;
                rs[y] = (a + c) + 2.0 * b;
            }
            gxs[t] = (rd[0] + rd[2]) + 2.0 * rd[1];
            gys[t] = rs[2] - rs[0];
            gts[t] = (rs[0] + rs[2]) + 2.0 * rs[1];
        }
        const double gx = (gxs[0] + gxs[2]) + 2.0 * gxs[1];
        const double gy = (gys[0] + gys[2]) + 2.0 * gys[1];
        const double gt = gts[2] - gts[0];
        return sqrt(gx * gx + gy * gy + gt * gt);
    }
    // plain direction, fp64: the separable form when no tap is NaN, else the general form (which skips NaN taps).  Every
    // non-centre tap has a non-zero weight in at least one gradient, so a finite separable result proves that no such
    // tap is NaN and the 27 NaN tests are only made when the result is not finite; the centre tap is tested directly.
    __device__ __forceinline__ ST finish() {
        if (sizeof(ST) == 8 && dir == TF_RED_SOBEL) {
            const double r = finish_separable();
            if (isfinite(r) && !is_nan(v[13])) return (ST)r;
            // Some tap is NaN (or the sums overflowed).  nansum skips a NaN tap, i.e. its d_k = v_k - centre counts as 0:
            // the same as the tap having the centre's value, so the separable form is redone with NaN taps replaced by
            // the centre (54 selects instead of the 300-instruction general walk; ~40 % of the warps of a frame with
            // 0.05 % bad pixels have such a lane).  Infinite taps / a NaN centre keep the general form.
            bool any_nan = false, any_inf = false;
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                any_nan |= is_nan(v[k]);
                any_inf |= !is_fin(v[k]) && !is_nan(v[k]);
            }
            if (!any_nan) return (ST)r;
            if (!any_inf && !is_nan(v[13])) {
                const KT c = v[13];
#pragma unroll
                for (int k = 0; k < 27; ++k) v[k] = is_nan(v[k]) ? c : v[k];
                return (ST)separable(v);    // finite: every tap is
            }
        }
        return finish_general();
    }
};

// All nine (dx, dy) in {-1,0,1}^2 taps of one warped slab at once (linear interpolation): when the nine sampling
// positions quantise consistently (same fraction, integer parts one apart - the overwhelmingly common case) and the
// 4x4 footprint is inside the image, load the 16 texels once and evaluate every tap from registers with exactly the
// arithmetic of remap_at<TF_LINEAR>.  Returns false when the caller must fall back to per-tap sampling.
// limit of the quantised centre position (minus one pixel) for which the 4 x 4 footprint lies inside an axis of n pixels
__device__ __forceinline__ unsigned patch_limit(int n) { return (unsigned)(min(max(n - 3, 0), 32764) << 5); }

// step 1: quantised positions of the three offsets per axis (p = fl32(fl32(flow + offset) + grid), convolve.py:56-63);
// false unless they share the fraction, their integer parts are one apart and the 4 x 4 footprint is inside the image
__device__ __forceinline__ bool patch9_locate(unsigned lim_x, unsigned lim_y, float2 f, int x, int y, int& qx, int& qy) {
    const float xf = (float)x, yf = (float)y;
    qx = quantise_fast(__fadd_rn(f.x, xf));
    qy = quantise_fast(__fadd_rn(f.y, yf));
    const int qxm = quantise_fast(__fadd_rn(__fadd_rn(f.x, -1.f), xf)), qxp = quantise_fast(__fadd_rn(__fadd_rn(f.x, 1.f), xf));
    const int qym = quantise_fast(__fadd_rn(__fadd_rn(f.y, -1.f), yf)), qyp = quantise_fast(__fadd_rn(__fadd_rn(f.y, 1.f), yf));
    return qxm + 32 == qx && qxp - 32 == qx && qym + 32 == qy && qyp - 32 == qy &&
           (unsigned)(qx - 32) < lim_x && (unsigned)(qy - 32) < lim_y;
}
// step 2: the 16 texels
template <typename T>
__device__ __forceinline__ void patch9_load(const T* __restrict__ img, int W, int qx, int qy, T p[4][4]) {
    const T* row = img + (((qy >> 5) - 1) * W + ((qx >> 5) - 1));
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) p[r][c] = row[c];
        row += W;
    }
}
// step 3: the nine taps, each with exactly the arithmetic of remap_at<TF_LINEAR>
template <typename T>
__device__ __forceinline__ void patch9_blend(const T p[4][4], int qx, int qy, T out[9]) {
    const float wfx = (float)(qx & 31) * (1.0f / 32.0f), wfy = (float)(qy & 31) * (1.0f / 32.0f);
    const float wx0 = 1.f - wfx, wy0 = 1.f - wfy;
    const T w00 = (T)(wy0 * wx0), w01 = (T)(wy0 * wfx), w10 = (T)(wfy * wx0), w11 = (T)(wfy * wfx);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
            out[dy * 3 + dx] = add_rn(add_rn(add_rn(mul_rn(p[dy][dx], w00), mul_rn(p[dy][dx + 1], w01)),
                                             mul_rn(p[dy + 1][dx], w10)), mul_rn(p[dy + 1][dx + 1], w11));
}

template <typename T>
__device__ __forceinline__ bool linear_patch9(const T* __restrict__ img, int W, unsigned lim_x, unsigned lim_y, float2 f, int x,
                                              int y, T out[9]) {
    int qx, qy;
    if (!patch9_locate(lim_x, lim_y, f, x, y, qx, qy)) return false;
    T p[4][4];
    patch9_load<T>(img, W, qx, qy, p);
    patch9_blend<T>(p, qx, qy, out);
    return true;
}

// Structures the reference and its callers actually use get kernels with the 3x3x3 mask as a compile-time constant
// (SB != 0): the 27-tap walk then unrolls to exactly the taps present.
constexpr unsigned SB_T3 = (1u << 4) | (1u << 13) | (1u << 22);                       // Flow.diff, filtered_tdiff
constexpr unsigned SB_CROSS7 = SB_T3 | (1u << 10) | (1u << 12) | (1u << 14) | (1u << 16);  // generate_binary_structure(3, 1)
constexpr unsigned SB_S5 = (1u << 10) | (1u << 12) | (1u << 13) | (1u << 14) | (1u << 16); // same-step cross
constexpr unsigned SB_L2 = (1u << 4) | (1u << 22);                                    // label.py:133-137
constexpr unsigned SB_FULL = (1u << 27) - 1u;                                           // sobel

#ifndef TF_GATHER_MINB
#define TF_GATHER_MINB 4
#endif
// one output pixel (x, y) of frame t: every interpolation, reducer, dtype and border case
template <typename SrcT, typename ST, int INTERP, int RC, unsigned SB>
__device__ __forceinline__ void gather_pixel(const GatherArgs& a, int x, int y, int t) {
    const int H = a.H, W = a.W;
    const long long hw = (long long)H * W;
    const int pix = y * W + x;
    const SrcT* cur = reinterpret_cast<const SrcT*>(a.cur0) + (long long)t * hw;
    const SrcT fill_s = fill_cast<SrcT>(a.fill);
    const ST fill_st = fill_cast<ST>(a.fill);
    const FrameSrc<SrcT> prev{(t > 0 || a.has_prev) ? cur - hw : nullptr, fill_s, H, W};
    const FrameSrc<SrcT> next{(t < a.n_frames - 1 || a.has_next) ? cur + hw : nullptr, fill_s, H, W};
    const unsigned structure = SB ? SB : a.structure;
    float2 bf = make_float2(0.f, 0.f), ff = make_float2(0.f, 0.f);
    if (structure & 0x1ffu) bf = __ldg(a.bflow0 + (long long)t * hw + pix);
    if (structure & (0x1ffu << 18)) ff = __ldg(a.fflow0 + (long long)t * hw + pix);
    const SrcT centre_src = cur[pix];

    RedNone<ST> rn{reinterpret_cast<ST*>(a.out) + (long long)t * hw, a.out_tap_stride, pix};
    RedDiff<ST> rd;
    RedStat<ST> rs;
    typedef typename std::conditional<(sizeof(SrcT) < sizeof(ST)), SrcT, ST>::type KeepT;
    RedSobel<ST, KeepT> rsob;
    if (RC == RC_DIFF) { rd.x[0] = rd.x[1] = rd.x[2] = (ST)0; }
    if (RC == RC_STAT) rs.init(a.reducer);
    if (RC == RC_SOBEL) rsob.init(a.reducer, cast_to<SrcT, ST>(centre_src));

    int n = 0;
#pragma unroll
    for (int slab = 0; slab < 3; ++slab) {
        const unsigned bits = (structure >> (9 * slab)) & 0x1ffu;
        if (!bits) continue;
        const float2 f = slab == 0 ? bf : ff;
        const FrameSrc<SrcT>& src = slab == 0 ? prev : next;
        // the (up to) nine tap values of this slab, gathered first so that the patch / per-tap choice is one branch
        SrcT tapv[9];
        bool oob1[9];
        if (slab == 1) {
            if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
                const SrcT* c0 = cur + pix - W - 1;      // interior: no tap leaves the image
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    if (!((bits >> j) & 1u)) continue;
                    oob1[j] = false;
                    tapv[j] = c0[(j / 3) * W + (j % 3)];
                }
            } else {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    if (!((bits >> j) & 1u)) continue;
                    const int yy = y + j / 3 - 1, xx = x + j % 3 - 1;
                    oob1[j] = !((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W);
                    tapv[j] = oob1[j] ? fill_s : cur[yy * W + xx];
                }
            }
        } else {
            bool have_patch = false;
            if constexpr (INTERP == TF_LINEAR && !std::is_same<SrcT, int>::value) {
                if (bits == 0x1ffu && src.p != nullptr) have_patch = linear_patch9<SrcT>(src.p, W, patch_limit(W), patch_limit(H), f, x, y, tapv);
            }
            if (!have_patch) {
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    if (!((bits >> j) & 1u)) continue;
                    // p = fl32(fl32(flow + offset) + grid)   (convolve.py:56-63)
                    const float px = __fadd_rn(__fadd_rn(f.x, (float)(j % 3 - 1)), (float)x);
                    const float py = __fadd_rn(__fadd_rn(f.y, (float)(j / 3 - 1)), (float)y);
                    tapv[j] = remap_at<INTERP, SrcT>(src, px, py);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (!((bits >> j) & 1u)) continue;
            // same-step taps outside the image take the fill value in the STACK dtype (convolve.py:140-142)
            const ST v = (slab == 1 && oob1[j]) ? fill_st : cast_to<SrcT, ST>(tapv[j]);
            const int k = slab * 9 + j;
            if (RC == RC_NONE) rn.add(n, k, v);
            if (RC == RC_DIFF) rd.add(n, k, v);
            if (RC == RC_STAT) rs.add(n, k, v);
            if (RC == RC_SOBEL) rsob.add(n, k, (KeepT)v);
            ++n;
        }
    }
    if (RC != RC_NONE) {
        ST r;
        if (RC == RC_DIFF) r = rd.finish();
        if (RC == RC_STAT) r = rs.finish();
        if (RC == RC_SOBEL) r = rsob.finish();
        if (is_nan(centre_src)) r = fill_st;  // res[np.isnan(data)] = fill_value   (convolve.py:346-347)
        reinterpret_cast<ST*>(a.out)[(long long)t * hw + pix] = r;
    }
}

template <typename SrcT, typename ST, int INTERP, int RC, unsigned SB>
__global__ void __launch_bounds__(256, (RC == RC_SOBEL ? TF_GATHER_MINB : 1)) sl_gather_kernel(GatherArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    gather_pixel<SrcT, ST, INTERP, RC, SB>(a, x, y, blockIdx.z);
}

// ------------------------------------------------------------------------------------------------------------------
// Lean kernels for the two other operators the reference runs on every frame: Flow.diff (flow.py:182-186; taps t-1, t, t+1
// at the centre, RedDiff) and Flow.convolve with its default structure (generate_binary_structure(3, 1), func = None:
// the seven taps stacked), fp32, linear interpolation.  The 27-tap kernel above is bound by its two dependent memory
// round trips (flow, then texels; ncu: 11-18 long-scoreboard stalls per issue at 39-41 % of the DRAM bandwidth).  Here a
// thread walks LEAN_ROWS rows (8 apart, so that the CTA's warps stay on adjacent rows) and requests the flow vectors of
// its next row before it samples the current one: one exposed round trip per pixel.  Samples whose 2 x 2 footprint is
// not inside the image, and the first / last frame of a series, go through gather_pixel (out of line).  In the diff kernel the
// flow vectors (read once) and the result (written once) use streaming loads / stores, so that they do not push the frames
// -- which the passes over frames t - 1, t and t + 1 all read -- out of L2 (0.563 -> 0.541 ms per 24 CONUS frames; the same
// hints make the seven-plane stack slower, 0.96 -> 1.06 ms, so it keeps the default policy).
// ------------------------------------------------------------------------------------------------------------------
template <typename SrcT, typename ST, int INTERP, int RC, unsigned SB>
__device__ __noinline__ void gather_pixel_call(const void* cur0, const float2* fflow0, const float2* bflow0, void* out,
                                               long long out_tap_stride, int n_frames, int has_prev, int has_next, int H,
                                               int W, int reducer, double fill, int x, int y, int t) {
    GatherArgs a;
    a.cur0 = cur0; a.fflow0 = fflow0; a.bflow0 = bflow0; a.out = out; a.out_tap_stride = out_tap_stride;
    a.n_frames = n_frames; a.has_prev = has_prev; a.has_next = has_next; a.H = H; a.W = W; a.structure = SB;
    a.reducer = reducer; a.fill = fill;
    gather_pixel<SrcT, ST, INTERP, RC, SB>(a, x, y, t);
}

// one bilinear sample whose footprint is known to be inside the image: remap_at<TF_LINEAR>'s arithmetic
__device__ __forceinline__ float lean_sample(const float* __restrict__ img, int W, int qx, int qy) {
    const float* p = img + ((qy >> 5) * W + (qx >> 5));
    const float v00 = __ldg(p), v01 = __ldg(p + 1), v10 = __ldg(p + W), v11 = __ldg(p + W + 1);
    const float fx = (float)(qx & 31) * (1.0f / 32.0f), fy = (float)(qy & 31) * (1.0f / 32.0f);
    const float wx0 = 1.f - fx, wy0 = 1.f - fy;
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, wy0 * wx0), __fmul_rn(v01, wy0 * fx)), __fmul_rn(v10, fy * wx0)),
                     __fmul_rn(v11, fy * fx));
}

constexpr int LEAN_ROWS = 4;
enum { LEAN_DIFF = 0, LEAN_CROSS7 = 1 };
template <int MODE> __device__ __forceinline__ float2 ld_flow(const float2* p) { return MODE == LEAN_DIFF ? __ldcs(p) : __ldg(p); }

// (measured, ms per 24 CONUS frames: diff 0.558 at 8 CTAs per SM / 0.591 at 6; cross-7 1.058 / 0.953; 27-tap kernel 0.879 / 1.277)
template <int MODE>
__global__ void __launch_bounds__(256, MODE == LEAN_DIFF ? 8 : 6) sl_lean_kernel(GatherArgs a) {
    constexpr int RC = MODE == LEAN_DIFF ? RC_DIFF : RC_NONE;
    constexpr unsigned SB = MODE == LEAN_DIFF ? SB_T3 : SB_CROSS7;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int t = blockIdx.z;
    const int H = a.H, W = a.W;
    if (x >= W) return;
    const long long hw = (long long)H * W;
    const float* cur = reinterpret_cast<const float*>(a.cur0) + (long long)t * hw;
    const float2* bfl = a.bflow0 + (long long)t * hw;
    const float2* ffl = a.fflow0 + (long long)t * hw;
    float* out = reinterpret_cast<float*>(a.out) + (long long)t * hw;
    const bool series_inner = (t > 0 || a.has_prev) && (t < a.n_frames - 1 || a.has_next);
    const unsigned lim_x = (unsigned)(min(max(W - 1, 0), 32767) << 5), lim_y = (unsigned)(min(max(H - 1, 0), 32767) << 5);
    const float xf = (float)x;
    int y = blockIdx.y * (8 * LEAN_ROWS) + threadIdx.y;
    float2 bf_n = make_float2(0.f, 0.f), ff_n = bf_n;
    if (y < H && series_inner) { bf_n = ld_flow<MODE>(bfl + y * W + x); ff_n = ld_flow<MODE>(ffl + y * W + x); }
#pragma unroll 1
    for (int k = 0; k < LEAN_ROWS && y < H; ++k, y += 8) {
        const float2 bf = bf_n, ff = ff_n;
        if (k + 1 < LEAN_ROWS && y + 8 < H && series_inner) { bf_n = ld_flow<MODE>(bfl + (y + 8) * W + x); ff_n = ld_flow<MODE>(ffl + (y + 8) * W + x); }
        const int pix = y * W + x;
        const float yf = (float)y;
        // p = fl32(fl32(flow + 0) + grid) = fl32(flow + grid)   (convolve.py:56-63)
        const int qx0 = quantise_fast(__fadd_rn(bf.x, xf)), qy0 = quantise_fast(__fadd_rn(bf.y, yf));
        const int qx2 = quantise_fast(__fadd_rn(ff.x, xf)), qy2 = quantise_fast(__fadd_rn(ff.y, yf));
        bool fast = series_inner && (unsigned)qx0 < lim_x && (unsigned)qy0 < lim_y && (unsigned)qx2 < lim_x && (unsigned)qy2 < lim_y;
        if (MODE == LEAN_CROSS7) fast = fast && x >= 1 && x < W - 1 && y >= 1 && y < H - 1;
        if (fast) {
            const float v0 = lean_sample(cur - hw, W, qx0, qy0);
            const float v2 = lean_sample(cur + hw, W, qx2, qy2);
            const float c = __ldg(cur + pix);
            if (MODE == LEAN_DIFF) {
                RedDiff<float> rd;
                rd.x[0] = v0; rd.x[1] = c; rd.x[2] = v2;
                float r = rd.finish();
                if (c != c) r = (float)a.fill;            // res[np.isnan(data)] = fill_value   (convolve.py:346-347)
                __stcs(out + pix, r);
            } else {
                const long long ts = a.out_tap_stride;
                float* o = out + pix;
                o[0] = v0;
                o[ts] = __ldg(cur + pix - W);
                o[2 * ts] = __ldg(cur + pix - 1);
                o[3 * ts] = c;
                o[4 * ts] = __ldg(cur + pix + 1);
                o[5 * ts] = __ldg(cur + pix + W);
                o[6 * ts] = v2;
            }
        } else {
            gather_pixel_call<float, float, TF_LINEAR, RC, SB>(a.cur0, a.fflow0, a.bflow0, a.out, a.out_tap_stride, a.n_frames,
                                                               a.has_prev, a.has_next, H, W, a.reducer, a.fill, x, y, t);
        }
    }
}

template <int MODE>
static int launch_lean(const GatherArgs& a, cudaStream_t s) {
    const double per_px = 3.0 * 4 + 16.0 + (MODE == LEAN_DIFF ? 4.0 : 28.0);
    LaunchTimer lt(KC_GATHER, per_px * a.H * a.W * a.n_frames, s, cdiv(a.n_frames, 65535));
    const long long hw = (long long)a.H * a.W;
    for (int t0 = 0; t0 < a.n_frames; t0 += 65535) {
        GatherArgs b = a;
        const int nt = min(a.n_frames - t0, 65535);
        b.cur0 = reinterpret_cast<const float*>(a.cur0) + t0 * hw;
        b.fflow0 = a.fflow0 + t0 * hw;
        b.bflow0 = a.bflow0 + t0 * hw;
        b.out = reinterpret_cast<float*>(a.out) + t0 * hw;
        b.n_frames = nt;
        b.has_prev = (t0 > 0) ? 1 : a.has_prev;
        b.has_next = (t0 + nt < a.n_frames) ? 1 : a.has_next;
        sl_lean_kernel<MODE><<<dim3(cdiv(a.W, 32), cdiv(a.H, 8 * LEAN_ROWS), nt), dim3(32, 8), 0, s>>>(b);
    }
    return check_launch("tf_sl_convolve (lean)");
}

// ------------------------------------------------------------------------------------------------------------------
// Flow.sobel as the reference calls it (fp32 frames, fp64 stack, linear interpolation, plain direction; sobel.py:7-86
// through flow.py:160-180): the 27-tap kernel above spends 830 instructions per pixel, most of them on cases this call
// never meets.  Here a pixel whose three 3 x 3 slabs lie inside the series and the image
//   * quantises the warped positions without conversions (quantise_fast) and blends each warped slab from its 4 x 4 patch,
//   * reduces slab by slab: the separable (1,2,1) x (1,2,1) x (-1,0,1) sums of a slab are folded into the three gradient
//     accumulators as soon as its nine taps exist, in exactly the order RedSobel::separable adds them (bit-identical
//     results), so only nine taps are live at a time;
//   * handles NaN taps per slab (a slab's all-positive sum is finite iff its taps are): NaN -> the centre value, as nansum
//     of (tap - centre) does; infinities and everything at the image / series border go to gather_pixel.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sobel_slab_sums(const float v[9], double& gxs, double& gys, double& gts) {
    double rd[3], rs[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double p = (double)v[3 * r], q = (double)v[3 * r + 1], s = (double)v[3 * r + 2];
        rd[r] = s - p;
        rs[r] = (p + s) + 2.0 * q;
    }
    gxs = (rd[0] + rd[2]) + 2.0 * rd[1];
    gys = rs[2] - rs[0];
    gts = (rs[0] + rs[2]) + 2.0 * rs[1];
}

// NaN taps of a slab -> the centre value; false when a tap is infinite (general form needed)
__device__ __forceinline__ bool sobel_patch_nans(float v[9], float centre, int skip) {
    bool inf = false;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k == skip) continue;
        inf |= fabsf(v[k]) == INFINITY;
        v[k] = v[k] != v[k] ? centre : v[k];
    }
    return !inf;
}

#ifndef TF_SOBEL_MINB
#define TF_SOBEL_MINB 8
#endif
// rows per thread: 1 (measured, ms per 24 CONUS frames: 2.07 with one row, 2.32-2.35 with 2 or 4 rows and the next row's
// flow vectors requested ahead -- the loop-carried state spills at 64 registers, and 80 registers cost a CTA per SM)
#ifndef TF_SOBEL_ROWS
#define TF_SOBEL_ROWS 1
#endif
// block = 32 x TF_SOBEL_BY pixels (measured, ms per 24 CONUS frames: 32 x 4 at 8 CTAs per SM 2.00, 32 x 8 at 4 CTAs 2.07,
// 32 x 4 at 9 / 10 CTAs (56 / 48 registers, spills) 2.02 / 2.06)
#ifndef TF_SOBEL_BY
#define TF_SOBEL_BY 4
#endif
__global__ void __launch_bounds__(32 * TF_SOBEL_BY, TF_SOBEL_MINB) sobel_lin_f64_kernel(GatherArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int t = blockIdx.z;
    const int H = a.H, W = a.W;
    if (x >= W) return;
    const long long hw = (long long)H * W;
    const float* cur = reinterpret_cast<const float*>(a.cur0) + (long long)t * hw;
    const float2* bfl = a.bflow0 + (long long)t * hw;
    const float2* ffl = a.fflow0 + (long long)t * hw;
    const bool inner_xt = x >= 1 && x < W - 1 && (t > 0 || a.has_prev) && (t < a.n_frames - 1 || a.has_next);
    const unsigned lim_x = patch_limit(W), lim_y = patch_limit(H);
    // a thread walks TF_SOBEL_ROWS rows, 8 apart, and requests the flow vectors of its next row before it works on the
    // current one (they are the head of the pixel's dependent load chain)
    int y = blockIdx.y * (TF_SOBEL_BY * TF_SOBEL_ROWS) + threadIdx.y;
    float2 bf_n = make_float2(0.f, 0.f), ff_n = bf_n;
    if (y < H && inner_xt) { bf_n = __ldg(bfl + y * W + x); ff_n = __ldg(ffl + y * W + x); }
#pragma unroll 1
    for (int k = 0; k < TF_SOBEL_ROWS && y < H; ++k, y += TF_SOBEL_BY) {
    const float2 bf = bf_n, ff = ff_n;
    if (k + 1 < TF_SOBEL_ROWS && y + TF_SOBEL_BY < H && inner_xt) { bf_n = __ldg(bfl + (y + TF_SOBEL_BY) * W + x); ff_n = __ldg(ffl + (y + TF_SOBEL_BY) * W + x); }
    const int pix = y * W + x;
    double* out = reinterpret_cast<double*>(a.out) + (long long)t * hw + pix;
    const bool inner = inner_xt && y >= 1 && y < H - 1;
    bool done = false;
    if (inner) {
        float v[9];
        const float* c0 = cur + pix - W - 1;
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = c0[(j / 3) * W + (j % 3)];
        const float centre = v[4];
        if (centre != centre) {                  // res[np.isnan(data)] = fill_value   (convolve.py:346-347)
            *out = a.fill;
            continue;
        }
        // both warped patches are located and requested before anything is reduced: two dependent memory round trips per
        // pixel (flow, then texels) instead of three -- the kernel is bound by that latency chain, not by instruction issue
        int qx0, qy0, qx2, qy2;
        bool ok = fabsf(centre) != INFINITY;
        ok = patch9_locate(lim_x, lim_y, bf, x, y, qx0, qy0) && ok;
        ok = patch9_locate(lim_x, lim_y, ff, x, y, qx2, qy2) && ok;
        if (ok) {
            float p0[4][4], p2[4][4];
            patch9_load<float>(cur - hw, W, qx0, qy0, p0);
            patch9_load<float>(cur + hw, W, qx2, qy2, p2);
            double gx1, gy1, gt1;
            sobel_slab_sums(v, gx1, gy1, gt1);
            // the centre tap has no weight in any gradient: gx1 and gy1 together cover the other eight taps
            if (!(isfinite(gx1) && isfinite(gy1))) {
                ok = sobel_patch_nans(v, centre, 4);
                sobel_slab_sums(v, gx1, gy1, gt1);
            }
            double gx0, gy0, gt0, gx2, gy2, gt2;
            patch9_blend<float>(p0, qx0, qy0, v);
            sobel_slab_sums(v, gx0, gy0, gt0);
            if (!isfinite(gt0)) {
                ok = sobel_patch_nans(v, centre, -1) && ok;
                sobel_slab_sums(v, gx0, gy0, gt0);
            }
            patch9_blend<float>(p2, qx2, qy2, v);
            sobel_slab_sums(v, gx2, gy2, gt2);
            if (!isfinite(gt2)) {
                ok = sobel_patch_nans(v, centre, -1) && ok;
                sobel_slab_sums(v, gx2, gy2, gt2);
            }
            if (ok) {
                const double gx = (gx0 + gx2) + 2.0 * gx1;
                const double gy = (gy0 + gy2) + 2.0 * gy1;
                const double gt = gt2 - gt0;
                *out = sqrt(gx * gx + gy * gy + gt * gt);
                done = true;
            }
        }
    }
    if (!done)
        gather_pixel_call<float, double, TF_LINEAR, RC_SOBEL, SB_FULL>(a.cur0, a.fflow0, a.bflow0, a.out, a.out_tap_stride, a.n_frames,
                                                                       a.has_prev, a.has_next, H, W, a.reducer, a.fill, x, y, t);
    }
}

static thread_local double g_gather_bytes = 0;

template <typename SrcT, typename ST, int INTERP, int RC, unsigned SB = 0u>
static int launch_gather(const GatherArgs& a, cudaStream_t s) {
    dim3 block(32, 8);
    {
        int n_out = 1;
        if (RC == RC_NONE) { n_out = 0; for (int k = 0; k < 27; ++k) n_out += (a.structure >> k) & 1; }
        const double per_px = 3.0 * sizeof(SrcT) + 16.0 + (double)n_out * sizeof(ST);
        g_gather_bytes = per_px * a.H * a.W * a.n_frames;
    }
    LaunchTimer lt(KC_GATHER, g_gather_bytes, s, cdiv(a.n_frames, 65535));
    for (int t0 = 0; t0 < a.n_frames; t0 += 65535) {
        GatherArgs b = a;
        const int nt = min(a.n_frames - t0, 65535);
        const long long hw = (long long)a.H * a.W;
        b.cur0 = reinterpret_cast<const SrcT*>(a.cur0) + t0 * hw;
        b.fflow0 = a.fflow0 + t0 * hw;
        b.bflow0 = a.bflow0 + t0 * hw;
        b.out = reinterpret_cast<ST*>(a.out) + t0 * hw;
        b.n_frames = nt;
        b.has_prev = (t0 > 0) ? 1 : a.has_prev;
        b.has_next = (t0 + nt < a.n_frames) ? 1 : a.has_next;
        dim3 grid(cdiv(a.W, 32), cdiv(a.H, 8), nt);
        sl_gather_kernel<SrcT, ST, INTERP, RC, SB><<<grid, block, 0, s>>>(b);
    }
    return check_launch("tf_sl_convolve");
}

static int launch_sobel_fast(const GatherArgs& a, cudaStream_t s) {
    LaunchTimer lt(KC_GATHER, (3.0 * 4 + 16.0 + 8.0) * a.H * a.W * a.n_frames, s, cdiv(a.n_frames, 65535));
    const long long hw = (long long)a.H * a.W;
    for (int t0 = 0; t0 < a.n_frames; t0 += 65535) {
        GatherArgs b = a;
        const int nt = min(a.n_frames - t0, 65535);
        b.cur0 = reinterpret_cast<const float*>(a.cur0) + t0 * hw;
        b.fflow0 = a.fflow0 + t0 * hw;
        b.bflow0 = a.bflow0 + t0 * hw;
        b.out = reinterpret_cast<double*>(a.out) + t0 * hw;
        b.n_frames = nt;
        b.has_prev = (t0 > 0) ? 1 : a.has_prev;
        b.has_next = (t0 + nt < a.n_frames) ? 1 : a.has_next;
        sobel_lin_f64_kernel<<<dim3(cdiv(a.W, 32), cdiv(a.H, TF_SOBEL_BY * TF_SOBEL_ROWS), nt), dim3(32, TF_SOBEL_BY), 0, s>>>(b);
    }
    return check_launch("tf_sl_convolve (sobel)");
}

template <typename SrcT, typename ST, int INTERP>
static int dispatch_rc(const GatherArgs& a, int rc, cudaStream_t s) {
    switch (rc) {
        case RC_NONE: return launch_gather<SrcT, ST, INTERP, RC_NONE>(a, s);
        case RC_DIFF: return launch_gather<SrcT, ST, INTERP, RC_DIFF>(a, s);
        case RC_STAT: return launch_gather<SrcT, ST, INTERP, RC_STAT>(a, s);
        default: return launch_gather<SrcT, ST, INTERP, RC_SOBEL>(a, s);
    }
}

template <typename SrcT, typename ST>
static int dispatch_interp(const GatherArgs& a, int interp, int rc, cudaStream_t s) {
    switch (interp) {
        case TF_NEAREST: return dispatch_rc<SrcT, ST, TF_NEAREST>(a, rc, s);
        case TF_LINEAR: return dispatch_rc<SrcT, ST, TF_LINEAR>(a, rc, s);
        case TF_LANCZOS4: return dispatch_rc<SrcT, ST, TF_LANCZOS4>(a, rc, s);
        default: return dispatch_rc<SrcT, ST, TF_CUBIC>(a, rc, s);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K9 smooth_flow_step (flow.py:530-568):  a' = nanmean([a, -warp(b, by = a)])
// ------------------------------------------------------------------------------------------------------------------
template <int INTERP>
__global__ void __launch_bounds__(256) smooth_flow_kernel(const float* __restrict__ fwd, const float* __restrict__ bwd,
                                                          float* __restrict__ fwd_out, float* __restrict__ bwd_out,
                                                          long long stride, int H, int W) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int p = blockIdx.z >> 1, which = blockIdx.z & 1;
    const float* A = (which ? bwd : fwd) + p * stride;   // field being smoothed
    const float* B = (which ? fwd : bwd) + p * stride;   // opposite field, warped by A
    float* O = (which ? bwd_out : fwd_out) + p * stride;
    const long long pix = (long long)y * W + x;
    const float ax = A[2 * pix], ay = A[2 * pix + 1];
    const float px = __fadd_rn(ax, (float)x), py = __fadd_rn(ay, (float)y);
    float o[2] = {ax, ay};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        FlowCompSrc src{B + c, NAN, H, W};
        const float wv = -remap_at<INTERP, float>(src, px, py);
        // np.nanmean([a, w], 0): NaN -> 0, fp32 sum, divide by the count through fp64
        const float a0 = o[c];
        float sum = (a0 != a0 ? 0.f : a0) + (wv != wv ? 0.f : wv);
        const int cnt = (a0 == a0) + (wv == wv);
        o[c] = (float)((double)sum / (double)cnt);
    }
    O[2 * pix] = o[0];
    O[2 * pix + 1] = o[1];
}

// K7: end rules + clamp (flow.py:425-426, 60-61)
__global__ void __launch_bounds__(256) flow_finalise_kernel(float* __restrict__ fwd, float* __restrict__ bwd, int T,
                                                            long long frame_elems, float max_value, int clamp_all,
                                                            int mirror_first, int mirror_last) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (i >= frame_elems) return;
    const long long o = (long long)t * frame_elems + i;
    float f = fwd[o], b = bwd[o];
    if (mirror_last && t == T - 1) f = -b;
    if (mirror_first && t == 0) b = -f;
    if (clamp_all && max_value > 0.f) {
        // np.minimum(np.maximum(v, -m), m) propagates NaN
        if (f == f) f = fminf(fmaxf(f, -max_value), max_value);
        if (b == b) b = fminf(fmaxf(b, -max_value), max_value);
    }
    fwd[o] = f;
    bwd[o] = b;
}

// interpolateLanczos4 for the 32 table fractions; uploaded once per device
static int upload_lanczos4_table() {
    static std::mutex mu;
    static unsigned long long done_mask = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("lanczos table: no CUDA device"); return TF_ERR_CUDA; }
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && ((done_mask >> dev) & 1ull)) return TF_OK;
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    const double pi = 3.1415926535897932384626433832795;     // CV_PI
    float tab[32 * 8];
    for (int i = 0; i < 32 * 8; ++i) tab[i] = 0.f;
    tab[3] = 1.f;                                             // x < FLT_EPSILON: the centre tap alone
    for (int i = 1; i < 32; ++i) {
        const float x = (float)i * (1.f / 32.f);
        float* co = tab + 8 * i;
        float sum = 0.f;
        const double y0 = -(x + 3) * pi * 0.25, s0 = sin(y0), c0 = cos(y0);
        for (int k = 0; k < 8; ++k) {
            const double y = -(x + 3 - k) * pi * 0.25;        // (x + 3 - k) in float, as in OpenCV
            co[k] = (float)((cs[k][0] * s0 + cs[k][1] * c0) / (y * y));
            sum += co[k];
        }
        sum = 1.f / sum;
        for (int k = 0; k < 8; ++k) co[k] *= sum;
    }
    if (cudaMemcpyToSymbol(c_lanczos4, tab, sizeof(tab)) != cudaSuccess) {
        set_error("lanczos table: cudaMemcpyToSymbol failed");
        return TF_ERR_CUDA;
    }
    if (dev < 64) done_mask |= 1ull << dev;
    return TF_OK;
}

}  // namespace tf

using namespace tf;

extern "C" int tf_sl_convolve(const void* cur0, int n_frames, int has_prev, int has_next, const float* fflow0,
                              const float* bflow0, void* out, long long out_tap_stride, int H, int W, int src_dtype,
                              int stack_dtype, int interp, int reducer, const uint8_t* structure27, double fill,
                              void* stream) {
    if (n_frames == 0) return TF_OK;
    if (!cur0 || !fflow0 || !bflow0 || !out || !structure27 || n_frames < 0 || H <= 0 || W <= 0) {
        set_error("tf_sl_convolve: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    if (interp < TF_NEAREST || interp > TF_LANCZOS4) { set_error("tf_sl_convolve: unknown interpolation %d", interp); return TF_ERR_INVALID_ARGUMENT; }
    if (interp == TF_LANCZOS4) { const int rc_l = upload_lanczos4_table(); if (rc_l != TF_OK) return rc_l; }
    if (reducer < TF_RED_NONE || reducer > TF_RED_NANMIN) { set_error("tf_sl_convolve: unknown reducer %d", reducer); return TF_ERR_INVALID_ARGUMENT; }
    GatherArgs a{};
    a.cur0 = cur0; a.fflow0 = reinterpret_cast<const float2*>(fflow0); a.bflow0 = reinterpret_cast<const float2*>(bflow0);
    a.out = out; a.out_tap_stride = out_tap_stride; a.n_frames = n_frames; a.has_prev = has_prev; a.has_next = has_next;
    a.H = H; a.W = W; a.reducer = reducer; a.fill = fill;
    int n_taps = 0;
    for (int k = 0; k < 27; ++k) if (structure27[k]) { a.structure |= 1u << k; ++n_taps; }
    int rc;
    switch (reducer) {
        case TF_RED_NONE: rc = RC_NONE; break;
        case TF_RED_DIFF: rc = RC_DIFF; break;
        case TF_RED_SOBEL: case TF_RED_SOBEL_UPHILL: case TF_RED_SOBEL_DOWNHILL: rc = RC_SOBEL; break;
        default: rc = RC_STAT;
    }
    if (rc == RC_DIFF && n_taps != 3) { set_error("tf_sl_convolve: the diff reducer needs exactly 3 taps"); return TF_ERR_INVALID_ARGUMENT; }
    if (rc == RC_SOBEL && n_taps != 27) { set_error("tf_sl_convolve: the sobel reducers need the full 27-tap structure"); return TF_ERR_INVALID_ARGUMENT; }
    if (n_taps == 0) { set_error("tf_sl_convolve: empty structure"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    if (src_dtype == TF_I32) {
        if (interp != TF_NEAREST) { set_error("tf_sl_convolve: integer operands support nearest interpolation only (as cv2.remap)"); return TF_ERR_UNSUPPORTED; }
        if (stack_dtype != TF_I32) { set_error("tf_sl_convolve: integer operands need an int32 result dtype"); return TF_ERR_UNSUPPORTED; }
        if (rc == RC_NONE && a.structure == SB_L2) return launch_gather<int, int, TF_NEAREST, RC_NONE, SB_L2>(a, s);
        if (rc == RC_STAT && reducer == TF_RED_ANY && a.structure == SB_T3)
            return launch_gather<int, int, TF_NEAREST, RC_STAT, SB_T3>(a, s);
        if (rc == RC_NONE) return launch_gather<int, int, TF_NEAREST, RC_NONE>(a, s);
        if (rc == RC_STAT && (reducer == TF_RED_ANY || reducer == TF_RED_NANMAX || reducer == TF_RED_NANMIN))
            return launch_gather<int, int, TF_NEAREST, RC_STAT>(a, s);
        set_error("tf_sl_convolve: reducer %d is not available for int32 operands", reducer);
        return TF_ERR_UNSUPPORTED;
    }
    // compile-time structures for the combinations on the reference's call paths
    const unsigned sb = a.structure;
    const bool f32 = src_dtype == TF_F32, f64s = src_dtype == TF_F64, st32 = stack_dtype == TF_F32, st64 = stack_dtype == TF_F64;
    if (interp == TF_LINEAR) {
        static const bool lean_off = getenv("TF_GATHER_GENERAL") != nullptr;       // A/B: the 27-tap kernel for diff / cross-7
        if (f32 && st32 && rc == RC_DIFF && sb == SB_T3 && !lean_off) return launch_lean<LEAN_DIFF>(a, s);
        if (f32 && st32 && rc == RC_NONE && sb == SB_CROSS7 && !lean_off) return launch_lean<LEAN_CROSS7>(a, s);
        if (f32 && st32 && rc == RC_DIFF && sb == SB_T3) return launch_gather<float, float, TF_LINEAR, RC_DIFF, SB_T3>(a, s);
        if (f32 && st32 && rc == RC_STAT && sb == SB_T3) return launch_gather<float, float, TF_LINEAR, RC_STAT, SB_T3>(a, s);
        if (f64s && st32 && rc == RC_STAT && sb == SB_T3) return launch_gather<double, float, TF_LINEAR, RC_STAT, SB_T3>(a, s);
        if (f32 && st32 && rc == RC_STAT && sb == SB_S5) return launch_gather<float, float, TF_LINEAR, RC_STAT, SB_S5>(a, s);
        static const bool sobel_general = getenv("TF_SOBEL_GENERAL") != nullptr;   // A/B: the 27-tap kernel for every pixel
        if (f32 && st64 && reducer == TF_RED_SOBEL && sb == SB_FULL && !sobel_general) return launch_sobel_fast(a, s);
        if (f32 && st64 && rc == RC_SOBEL && sb == SB_FULL) return launch_gather<float, double, TF_LINEAR, RC_SOBEL, SB_FULL>(a, s);
        if (f32 && st32 && rc == RC_SOBEL && sb == SB_FULL) return launch_gather<float, float, TF_LINEAR, RC_SOBEL, SB_FULL>(a, s);
        if (f32 && st32 && rc == RC_NONE && sb == SB_CROSS7) return launch_gather<float, float, TF_LINEAR, RC_NONE, SB_CROSS7>(a, s);
    }
    if (interp == TF_CUBIC && f32 && st64 && rc == RC_SOBEL && sb == SB_FULL)
        return launch_gather<float, double, TF_CUBIC, RC_SOBEL, SB_FULL>(a, s);
    if (src_dtype == TF_F32 && stack_dtype == TF_F32) return dispatch_interp<float, float>(a, interp, rc, s);
    if (src_dtype == TF_F32 && stack_dtype == TF_F64) return dispatch_interp<float, double>(a, interp, rc, s);
    if (src_dtype == TF_F64 && stack_dtype == TF_F32) return dispatch_interp<double, float>(a, interp, rc, s);
    if (src_dtype == TF_F64 && stack_dtype == TF_F64) return dispatch_interp<double, double>(a, interp, rc, s);
    set_error("tf_sl_convolve: unsupported dtype combination src=%d stack=%d", src_dtype, stack_dtype);
    return TF_ERR_UNSUPPORTED;
}

extern "C" int tf_smooth_flow_step(const float* fwd, const float* bwd, float* fwd_out, float* bwd_out, long long stride,
                                   int n_pairs, int H, int W, int interp, void* stream) {
    if (n_pairs == 0) return TF_OK;
    if (!fwd || !bwd || !fwd_out || !bwd_out || n_pairs < 0 || H <= 0 || W <= 0) {
        set_error("tf_smooth_flow_step: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    if (fwd == fwd_out || bwd == bwd_out) { set_error("tf_smooth_flow_step: outputs must not alias inputs"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(32, 8);
    LaunchTimer lt(KC_SMOOTH, 32.0 * H * W * n_pairs, s, cdiv(n_pairs, 32767));
    for (int p0 = 0; p0 < n_pairs; p0 += 32767) {
        const int np = min(n_pairs - p0, 32767);
        dim3 grid(cdiv(W, 32), cdiv(H, 8), 2 * np);
        const float* f = fwd + p0 * stride; const float* b = bwd + p0 * stride;
        float* fo = fwd_out + p0 * stride; float* bo = bwd_out + p0 * stride;
        switch (interp) {
            case TF_NEAREST: smooth_flow_kernel<TF_NEAREST><<<grid, block, 0, s>>>(f, b, fo, bo, stride, H, W); break;
            case TF_LINEAR: smooth_flow_kernel<TF_LINEAR><<<grid, block, 0, s>>>(f, b, fo, bo, stride, H, W); break;
            case TF_CUBIC: smooth_flow_kernel<TF_CUBIC><<<grid, block, 0, s>>>(f, b, fo, bo, stride, H, W); break;
            case TF_LANCZOS4:
                if (upload_lanczos4_table() != TF_OK) return TF_ERR_CUDA;
                smooth_flow_kernel<TF_LANCZOS4><<<grid, block, 0, s>>>(f, b, fo, bo, stride, H, W);
                break;
            default: set_error("tf_smooth_flow_step: unknown interpolation %d", interp); return TF_ERR_INVALID_ARGUMENT;
        }
    }
    return check_launch("tf_smooth_flow_step");
}

extern "C" int tf_flow_finalise(float* fwd, float* bwd, int T, int H, int W, float max_value, int clamp_all,
                                int mirror_first, int mirror_last, void* stream) {
    if (T == 0) return TF_OK;
    if (!fwd || !bwd || T < 0 || H <= 0 || W <= 0) { set_error("tf_flow_finalise: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    const long long fe = (long long)H * W * 2;
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_FINALISE, (clamp_all && max_value > 0.f) ? 4.0 * fe * T * 4 : 4.0 * fe * 4, s,
                   (clamp_all && max_value > 0.f) ? cdiv(T, 65535) : 2);
    if (clamp_all && max_value > 0.f) {
        for (int t0 = 0; t0 < T; t0 += 65535) {
            const int nt = min(T - t0, 65535);
            dim3 grid((unsigned)((fe + 255) / 256), nt);
            // mirror flags only apply to the global first / last frame
            flow_finalise_kernel<<<grid, 256, 0, s>>>(fwd + t0 * fe, bwd + t0 * fe, nt, fe, max_value, 1,
                                                      mirror_first && t0 == 0, mirror_last && t0 + nt == T);
        }
    } else {
        // only the two end frames change
        dim3 grid((unsigned)((fe + 255) / 256), 1);
        if (T == 1) {
            // a one-frame shard: frame 0 is both ends (either rule may apply alone, e.g. the last rank of a sharded run)
            if (mirror_first || mirror_last)
                flow_finalise_kernel<<<grid, 256, 0, s>>>(fwd, bwd, 1, fe, 0.f, 0, mirror_first, mirror_last);
        } else {
            if (mirror_first)
                flow_finalise_kernel<<<grid, 256, 0, s>>>(fwd, bwd, 2, fe, 0.f, 0, 1, 0);
            if (mirror_last)
                flow_finalise_kernel<<<grid, 256, 0, s>>>(fwd + (T - 1) * fe, bwd + (T - 1) * fe, 1, fe, 0.f, 0, 0, 1);
        }
    }
    return check_launch("tf_flow_finalise");
}
