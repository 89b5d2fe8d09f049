// K3: Farneback polynomial expansion (OpenCV FarnebackPolyExp, polyN = 5): separable 11-tap tile convolution with
// replicate borders.  Input level image I (h, w) fp32 -> per image a float4 plane (c0..c3) followed by a float plane
// (c4), so the iteration kernel's bilinear gathers are one 16-byte and one 4-byte coalesced load per tap.
// Shared-memory tile with a 5-pixel halo: vertical pass to three moment arrays, horizontal pass to six sums.
// HBM traffic: 4 B/px read (+ halo re-reads that hit L2) and 20 B/px written.
#include "farneback_internal.cuh"

namespace tf {

constexpr int PE_TW = 64, PE_TH = 32, PE_N = 5;
constexpr int PE_SW = PE_TW + 2 * PE_N;      // 74 columns incl. halo
constexpr int PE_SWP = PE_SW + 2;            // padded row pitch (76)
constexpr int PE_SH = PE_TH + 2 * PE_N;      // 42 rows incl. halo

__global__ void __launch_bounds__(256) polyexp_kernel(const float* __restrict__ I, float* __restrict__ R,
                                                      long long img_stride, int h, int w, PolyConsts pc) {
    __shared__ float s_in[PE_SH][PE_SWP];
    __shared__ float s_v[3][PE_TH][PE_SWP];
    const int img = blockIdx.z;
    const int x0 = blockIdx.x * PE_TW, y0 = blockIdx.y * PE_TH;
    const float* src = I + (long long)img * h * w;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < PE_SH * PE_SW; idx += 256) {
        const int r = idx / PE_SW, c = idx - r * PE_SW;
        const int gy = min(max(y0 + r - PE_N, 0), h - 1);
        const int gx = min(max(x0 + c - PE_N, 0), w - 1);
        s_in[r][c] = src[(long long)gy * w + gx];
    }
    __syncthreads();

    // vertical pass (fp32, same accumulation order as OpenCV: k = 1..n)
    for (int idx = tid; idx < PE_TH * PE_SW; idx += 256) {
        const int r = idx / PE_SW, c = idx - r * PE_SW;
        // replicate in y is relative to the IMAGE, which the clamped tile load already provides
        const float ctr = s_in[r + PE_N][c];
        float t0 = ctr * pc.g[0], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 1; k <= PE_N; ++k) {
            const float a = s_in[r + PE_N - k][c], b = s_in[r + PE_N + k][c];
            const float p = a + b;
            t0 = t0 + pc.g[k] * p;
            t1 = t1 + pc.xg[k] * (b - a);
            t2 = t2 + pc.xxg[k] * p;
        }
        s_v[0][r][c] = t0;
        s_v[1][r][c] = t1;
        s_v[2][r][c] = t2;
    }
    __syncthreads();

    // horizontal pass
    const int c = tid & (PE_TW - 1);
    const int gx = x0 + c;
    const long long plane = (long long)h * w;
    float4* dst_a = reinterpret_cast<float4*>(R + (long long)img * img_stride);
    float* dst_b = R + (long long)img * img_stride + 4 * plane;
    for (int r = tid >> 6; r < PE_TH; r += 4) {
        const int gy = y0 + r;
        if (gx >= w || gy >= h) continue;
        const float* v0 = &s_v[0][r][c + PE_N];
        const float* v1 = &s_v[1][r][c + PE_N];
        const float* v2 = &s_v[2][r][c + PE_N];
        float b1 = v0[0] * pc.g[0], b2 = 0.f, b3 = v1[0] * pc.g[0], b4 = 0.f, b5 = v2[0] * pc.g[0], b6 = 0.f;
#pragma unroll
        for (int k = 1; k <= PE_N; ++k) {
            const float tg = v0[k] + v0[-k];
            b1 += tg * pc.g[k];
            b4 += tg * pc.xxg[k];
            b2 += (v0[k] - v0[-k]) * pc.xg[k];
            b3 += (v1[k] + v1[-k]) * pc.g[k];
            b6 += (v1[k] - v1[-k]) * pc.xg[k];
            b5 += (v2[k] + v2[-k]) * pc.g[k];
        }
        const long long o = (long long)gy * w + gx;
        dst_a[o] = make_float4(b3 * pc.ig11, b2 * pc.ig11, b1 * pc.ig03 + b5 * pc.ig33, b1 * pc.ig03 + b4 * pc.ig33);
        dst_b[o] = b6 * pc.ig55;
    }
}

int launch_polyexp(const float* I, float* R, long long img_stride, int n_img, int h, int w, const PolyConsts& pc,
                   cudaStream_t s) {
    LaunchTimer lt(KC_POLYEXP, 24.0 * h * w * n_img, s, cdiv(n_img, 65535));
    for (int z0 = 0; z0 < n_img; z0 += 65535) {
        const int nz = min(n_img - z0, 65535);
        dim3 g(cdiv(w, PE_TW), cdiv(h, PE_TH), nz);
        polyexp_kernel<<<g, 256, 0, s>>>(I + (long long)z0 * h * w, R + (long long)z0 * img_stride, img_stride, h, w, pc);
    }
    return check_launch("polyexp");
}

// FarnebackPrepareGaussian: fp32 taps, fp64 moment matrix inverse (closed form for the sparse 6x6)
PolyConsts make_poly_consts(int n, double sigma) {
    PolyConsts pc{};
    if (sigma < 1.1920929e-07) sigma = n * 0.3;
    float g[11];
    double s = 0;
    for (int x = -n; x <= n; ++x) {
        g[x + n] = (float)exp(-x * x / (2 * sigma * sigma));
        s += g[x + n];
    }
    s = 1. / s;
    for (int x = -n; x <= n; ++x) g[x + n] = (float)(g[x + n] * s);
    for (int k = 0; k <= n; ++k) {
        pc.g[k] = g[n + k];
        pc.xg[k] = (float)(k * g[n + k]);
        pc.xxg[k] = (float)(k * k * g[n + k]);
    }
    double G00 = 0, G11 = 0, G33 = 0, G55 = 0;
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            const double gg = (double)g[y + n] * g[x + n];
            G00 += gg;
            G11 += gg * x * x;
            G33 += gg * x * x * x * x;
            G55 += gg * x * x * y * y;
        }
    // G = [[G00,0,0,G11,G11,0],[0,G11,..],[..G11..],[G11,0,0,G33,G55,0],[G11,0,0,G55,G33,0],[0,..,G55]]
    // inverse entries needed: (1,1), (0,3), (3,3), (5,5).  The {0,3,4} block is
    //   A = [[a, b, b], [b, c, d], [b, d, c]] with a=G00, b=G11, c=G33, d=G55.
    const double a = G00, b = G11, c = G33, d = G55;
    const double det = a * (c * c - d * d) - 2 * b * b * (c - d);
    pc.ig11 = (float)(1.0 / G11);
    pc.ig03 = (float)(-(b * (c - d)) / det);   // cofactor(3,0)/det = -(b*c - b*d)/det
    pc.ig33 = (float)((a * c - b * b) / det);
    pc.ig55 = (float)(1.0 / G55);
    return pc;
}

}  // namespace tf
