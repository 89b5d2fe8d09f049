// K2: one pyramid level = convertTo(CV_32F) -> GaussianBlur(ksize, sigma, REFLECT_101) -> resize(INTER_LINEAR),
// always from the full-resolution u8 image (OpenCV FarnebackOpticalFlow::calc, fastPyramids = false).
// K6: bilinear flow up-sampling (resize(prevFlow) * 1/pyrScale).
//
// Only the source rows/columns the bilinear resize actually reads are blurred: pass A blurs vertically at the
// (at most two) source rows of every destination row, pass B blurs those rows horizontally at the (at most two)
// source columns of every destination pixel and combines the four values with OpenCV's resize weights.
#include "tf_common.cuh"
#include "farneback_internal.cuh"

namespace tf {

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

// cv::resize INTER_LINEAR source coordinate: index of the first tap and fp32 weight of the second
__device__ __forceinline__ void resize_coord(int d, double scale, int src_n, int& i0, int& i1, float& f) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int ix = (int)floorf(fx);
    fx -= (float)ix;
    if (ix < 0) { fx = 0.f; ix = 0; }
    if (ix >= src_n - 1) { fx = 0.f; ix = src_n - 1; }
    i0 = ix;
    i1 = min(ix + 1, src_n - 1);
    f = fx;
}

// pass A.  tmp: (n_img, h*rp, W) fp32, rp = rows per destination row (1 when h == H, else 2)
__global__ void __launch_bounds__(256) blur_v_kernel(const uint8_t* __restrict__ q0, const uint8_t* __restrict__ q1,
                                                     float* __restrict__ tmp, int H, int W, int h, int rp,
                                                     double scale_y, BlurTaps taps) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;  // 2*pair + which
    if (x >= W) return;
    const uint8_t* src = ((img & 1) ? q1 : q0) + (long long)(img >> 1) * H * W;
    int sy;
    if (rp == 1) {
        sy = r;
    } else {
        int y0, y1; float fy;
        resize_coord(r >> 1, scale_y, H, y0, y1, fy);
        sy = (r & 1) ? y1 : y0;
    }
    const int rad = taps.ksize >> 1;
    float acc = 0.f;
    for (int k = 0; k < taps.ksize; ++k) {
        int yy = reflect101(sy + k - rad, H);
        acc += taps.w[k] * (float)src[(long long)yy * W + x];
    }
    tmp[((long long)img * (h * rp) + r) * W + x] = acc;
}

// exact u8 -> fp32 without the quarter-rate I2F: splice the byte into the mantissa of 2^23 and subtract 2^23
__device__ __forceinline__ float byte_to_float(unsigned v, unsigned sel) {
    return __uint_as_float(__byte_perm(v, 0x4B000000u, sel)) - 8388608.f;
}

// pass A, 4 columns per thread (W % 4 == 0): one 32-bit load per tap row, float4 store
__global__ void __launch_bounds__(256) blur_v4_kernel(const uint8_t* __restrict__ q0, const uint8_t* __restrict__ q1,
                                                      float* __restrict__ tmp, int H, int W, int h, int rp,
                                                      double scale_y, BlurTaps taps) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;   // group of 4 columns
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    const int W4 = W >> 2;
    if (x4 >= W4) return;
    const unsigned* src = reinterpret_cast<const unsigned*>(((img & 1) ? q1 : q0) + (long long)(img >> 1) * H * W);
    int sy;
    if (rp == 1) {
        sy = r;
    } else {
        int y0, y1; float fy;
        resize_coord(r >> 1, scale_y, H, y0, y1, fy);
        sy = (r & 1) ? y1 : y0;
    }
    const int rad = taps.ksize >> 1;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool interior = sy - rad >= 0 && sy + rad < H;
#pragma unroll 4
    for (int k = 0; k < taps.ksize; ++k) {
        const int yy = interior ? sy + k - rad : reflect101(sy + k - rad, H);
        const unsigned c = __ldg(src + yy * W4 + x4);
        const float wk = taps.w[k];
        acc.x += wk * byte_to_float(c, 0x7540u);
        acc.y += wk * byte_to_float(c, 0x7541u);
        acc.z += wk * byte_to_float(c, 0x7542u);
        acc.w += wk * byte_to_float(c, 0x7543u);
    }
    reinterpret_cast<float4*>(tmp + ((long long)img * (h * rp) + r) * W)[x4] = acc;
}

// pass B at the full-resolution level (w == W, no resize), 4 outputs per thread (W % 4 == 0)
__global__ void __launch_bounds__(256) blur_h4_fullres_kernel(const float* __restrict__ tmp, float* __restrict__ out, int W,
                                                              int h, BlurTaps taps) {
    const int x4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int img = blockIdx.z;
    const int W4 = W >> 2;
    if (x4 >= W4) return;
    const float* row = tmp + ((long long)img * h + j) * W;
    const int rad = taps.ksize >> 1;
    const int xb = 4 * x4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (taps.ksize == 3 && xb >= 1 && xb + 4 < W) {
        const float4 m = reinterpret_cast<const float4*>(row)[x4];
        const float l = row[xb - 1], rr = row[xb + 4];
        const float w0 = taps.w[0], w1 = taps.w[1], w2 = taps.w[2];
        acc.x = (w0 * l + w1 * m.x) + w2 * m.y;
        acc.y = (w0 * m.x + w1 * m.y) + w2 * m.z;
        acc.z = (w0 * m.y + w1 * m.z) + w2 * m.w;
        acc.w = (w0 * m.z + w1 * m.w) + w2 * rr;
    } else {
        for (int k = 0; k < taps.ksize; ++k) {
            const float wk = taps.w[k];
            acc.x += wk * row[reflect101(xb + k - rad, W)];
            acc.y += wk * row[reflect101(xb + 1 + k - rad, W)];
            acc.z += wk * row[reflect101(xb + 2 + k - rad, W)];
            acc.w += wk * row[reflect101(xb + 3 + k - rad, W)];
        }
    }
    reinterpret_cast<float4*>(out + ((long long)img * h + j) * W)[x4] = acc;
}

// pass B.  out: (n_img, h, w) fp32
__global__ void __launch_bounds__(256) blur_h_resize_kernel(const float* __restrict__ tmp, float* __restrict__ out, int W,
                                                            int h, int w, int rp, double scale_x, double scale_y,
                                                            int H, BlurTaps taps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int img = blockIdx.z;
    if (i >= w) return;
    const int rad = taps.ksize >> 1;
    const float* rows = tmp + ((long long)img * (h * rp) + (long long)j * rp) * W;
    float res;
    if (w == W && rp == 1) {
        float acc = 0.f;
        for (int k = 0; k < taps.ksize; ++k) acc += taps.w[k] * rows[reflect101(i + k - rad, W)];
        res = acc;
    } else {
        int x0, x1, y0, y1; float fx, fy;
        resize_coord(i, scale_x, W, x0, x1, fx);
        resize_coord(j, scale_y, H, y0, y1, fy);
        const float* row0 = rows;
        const float* row1 = rp == 2 ? rows + W : rows;
        float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
        if (x1 == x0 + 1 && x0 - rad >= 0 && x1 + rad < W) {
            // interior: the two column positions are neighbours, so one pass over ksize+1 values feeds both
            const float* p0 = row0 + x0 - rad;
            const float* p1 = row1 + x0 - rad;
            float a0 = p0[0], a1 = p1[0];
#pragma unroll 4
            for (int k = 0; k < taps.ksize; ++k) {
                const float wk = taps.w[k];
                const float n0 = p0[k + 1], n1 = p1[k + 1];
                b00 += wk * a0;
                b01 += wk * n0;
                b10 += wk * a1;
                b11 += wk * n1;
                a0 = n0;
                a1 = n1;
            }
        } else {
            for (int k = 0; k < taps.ksize; ++k) {
                const float wk = taps.w[k];
                const int xa = reflect101(x0 + k - rad, W), xb = reflect101(x1 + k - rad, W);
                b00 += wk * row0[xa];
                b01 += wk * row0[xb];
                b10 += wk * row1[xa];
                b11 += wk * row1[xb];
            }
        }
        const float r0 = b00 * (1.f - fx) + b01 * fx;
        const float r1 = b10 * (1.f - fx) + b11 * fx;
        res = r0 * (1.f - fy) + r1 * fy;
    }
    out[((long long)img * h + j) * w + i] = res;
}

// K6: dst (n_fields, h, w, 2) = resize_linear(src (n_fields, sh, sw, 2)) * mul ; zero-fill when src == nullptr
__global__ void __launch_bounds__(256) flow_upsample_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int sh,
                                                            int sw, int h, int w, double scale_x, double scale_y,
                                                            float mul) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int f = blockIdx.z;
    if (i >= w) return;
    float2 o = make_float2(0.f, 0.f);
    if (src != nullptr) {
        int x0, x1, y0, y1; float fx, fy;
        resize_coord(i, scale_x, sw, x0, x1, fx);
        resize_coord(j, scale_y, sh, y0, y1, fy);
        const float2* s = src + (long long)f * sh * sw;
        const float2 a = s[(long long)y0 * sw + x0], b = s[(long long)y0 * sw + x1];
        const float2 c = s[(long long)y1 * sw + x0], d = s[(long long)y1 * sw + x1];
        const float ax = a.x * (1.f - fx) + b.x * fx, ay = a.y * (1.f - fx) + b.y * fx;
        const float cx = c.x * (1.f - fx) + d.x * fx, cy = c.y * (1.f - fx) + d.y * fx;
        o.x = (ax * (1.f - fy) + cx * fy) * mul;
        o.y = (ay * (1.f - fy) + cy * fy) * mul;
    }
    dst[((long long)f * h + j) * w + i] = o;
}

int launch_pyramid_level(const uint8_t* q0, const uint8_t* q1, int n_pairs, int H, int W, int h, int w, int ksize,
                         double sigma, float* tmp, float* out, cudaStream_t s) {
    BlurTaps taps;
    taps.ksize = ksize;
    if (ksize > kMaxBlurTaps) { set_error("pyramid: blur kernel too wide (%d)", ksize); return TF_ERR_UNSUPPORTED; }
    if (sigma <= 0 && ksize == 3) {
        taps.w[0] = 0.25f; taps.w[1] = 0.5f; taps.w[2] = 0.25f;
    } else {
        if (sigma <= 0) sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8;
        double sum = 0, v[kMaxBlurTaps];
        for (int i = 0; i < ksize; ++i) {
            double x = i - (ksize - 1) * 0.5;
            v[i] = exp(-0.5 * x * x / (sigma * sigma));
            sum += v[i];
        }
        for (int i = 0; i < ksize; ++i) taps.w[i] = (float)(v[i] / sum);
    }
    const int rp = (h == H) ? 1 : 2;
    const double sx = (double)W / w, sy = (double)H / h;
    const int n_img = 2 * n_pairs;
    LaunchTimer lt(KC_PYRAMID, (2.0 * H * W + 8.0 * h * w) * n_pairs, s, 2 * cdiv(n_img, 65534));
    for (int z0 = 0; z0 < n_img; z0 += 65534) {
        const int nz = min(n_img - z0, 65534);  // even, so image parity is preserved
        const uint8_t* a0 = q0 + (long long)(z0 / 2) * H * W;
        const uint8_t* a1 = q1 + (long long)(z0 / 2) * H * W;
        float* tz = tmp + (long long)z0 * h * rp * W;
        const bool vec4 = (W % 4 == 0) && (((uintptr_t)a0 | (uintptr_t)a1) % 4 == 0) && ((uintptr_t)tz % 16 == 0);
        if (vec4) {
            dim3 ga(cdiv(W / 4, 128), h * rp, nz);
            blur_v4_kernel<<<ga, 128, 0, s>>>(a0, a1, tz, H, W, h, rp, sy, taps);
        } else {
            dim3 ga(cdiv(W, 256), h * rp, nz);
            blur_v_kernel<<<ga, 256, 0, s>>>(a0, a1, tz, H, W, h, rp, sy, taps);
        }
        float* oz = out + (long long)z0 * h * w;
        if (vec4 && w == W && rp == 1 && (uintptr_t)oz % 16 == 0) {
            dim3 gb(cdiv(W / 4, 128), h, nz);
            blur_h4_fullres_kernel<<<gb, 128, 0, s>>>(tz, oz, W, h, taps);
        } else {
            dim3 gb(cdiv(w, 256), h, nz);
            blur_h_resize_kernel<<<gb, 256, 0, s>>>(tz, oz, W, h, w, rp, sx, sy, H, taps);
        }
    }
    return check_launch("pyramid level");
}

int launch_flow_upsample(const float* src, float* dst, int n_fields, int sh, int sw, int h, int w, float mul,
                         cudaStream_t s) {
    const double sx = sw > 0 ? (double)sw / w : 1.0, sy = sh > 0 ? (double)sh / h : 1.0;
    LaunchTimer lt(KC_UPSAMPLE, (8.0 * h * w + 8.0 * sh * sw) * n_fields, s, cdiv(n_fields, 65535));
    for (int z0 = 0; z0 < n_fields; z0 += 65535) {
        const int nz = min(n_fields - z0, 65535);
        dim3 g(cdiv(w, 256), h, nz);
        flow_upsample_kernel<<<g, 256, 0, s>>>(
            src ? reinterpret_cast<const float2*>(src) + (long long)z0 * sh * sw : nullptr,
            reinterpret_cast<float2*>(dst) + (long long)z0 * h * w, sh, sw, h, w, sx, sy, mul);
    }
    return check_launch("flow upsample");
}

}  // namespace tf
