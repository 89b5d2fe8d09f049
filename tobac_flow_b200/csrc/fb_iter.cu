// K4+K5: one fused Farneback iteration = FarnebackUpdateMatrices + FarnebackUpdateFlow_Blur (box window).
//
// OpenCV's sweep is a pure Jacobi update (the matrices it refreshes behind the sliding window are never re-read in the
// same sweep), so flow buffers ping-pong and M is never materialised in HBM:
//   phase 1  M(y, x) for the output tile + 6-px halo (replicate-clamped coordinates): flow, R0, bilinear R1 at p+flow
//   phase 2  vertical 13-sums, sliding, in place in shared memory (one thread per column-channel)
//   phase 3  horizontal 13-sums (sliding over 8 outputs per thread), 2x2 solve in registers, float2 store
// Algorithmic HBM bytes per pixel-iteration: flow 8 + R0 20 + R1 20 read, flow 8 written = 56 B.
#include "farneback_internal.cuh"

namespace tf {

constexpr int IT_TW = 64, IT_TH = 32;

template <int HALO>
struct IterCfg {
    static constexpr int RW = IT_TW + 2 * HALO;   // region width  (76)
    static constexpr int RH = IT_TH + 2 * HALO;   // region height (44)
    static constexpr int PITCH = (RW + 3) / 4 * 4;
    static constexpr int SMEM_BYTES = 5 * RH * PITCH * (int)sizeof(float);
};

__device__ __forceinline__ float border_factor(int p, int n) {
    // border[] = {0.14, 0.14, 0.4472, 0.4472, 0.4472} applied from both sides
    float s = 1.f;
    if (p < 5) s *= (p < 2 ? 0.14f : 0.4472f);
    const int q = n - 1 - p;
    if (q < 5) s *= (q < 2 ? 0.14f : 0.4472f);
    return s;
}

// FarnebackUpdateMatrices for one pixel.  R planes: R[c * plane + y * w + x]
__device__ __forceinline__ void update_matrix_px(const float* __restrict__ R0, const float* __restrict__ R1,
                                                 long long plane, int w, int h, int x, int y, float dx, float dy,
                                                 float m[5]) {
    const long long o = (long long)y * w + x;
    float fx = (float)x + dx, fy = (float)y + dy;
    const float flx = floorf(fx), fly = floorf(fy);
    const int x1 = (int)flx, y1 = (int)fly;
    fx -= flx;
    fy -= fly;
    float r2, r3, r4, r5, r6;
    const float c2 = R0[2 * plane + o], c3 = R0[3 * plane + o], c4 = R0[4 * plane + o];
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const float* p = R1 + (long long)y1 * w + x1;
#define TF_BILIN(c) (a00 * p[(c) * plane] + a01 * p[(c) * plane + 1] + a10 * p[(c) * plane + w] + a11 * p[(c) * plane + w + 1])
        r2 = TF_BILIN(0);
        r3 = TF_BILIN(1);
        r4 = TF_BILIN(2);
        r5 = TF_BILIN(3);
        r6 = TF_BILIN(4);
#undef TF_BILIN
        r4 = (c2 + r4) * 0.5f;
        r5 = (c3 + r5) * 0.5f;
        r6 = (c4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = c2;
        r5 = c3;
        r6 = c4 * 0.5f;
    }
    r2 = (R0[o] - r2) * 0.5f;
    r3 = (R0[plane + o] - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float sc = border_factor(x, w) * border_factor(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

// a*b - c*d with one rounding error in the result (Kahan)
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    const float cd = c * d;
    const float err = fmaf(-c, d, cd);
    const float dop = fmaf(a, b, -cd);
    return dop + err;
}

template <int HALO>
__global__ void __launch_bounds__(256) fb_iter_kernel(const float* __restrict__ R, const float* __restrict__ flow_in,
                                                      float* __restrict__ out_fwd, long long fwd_stride,
                                                      float* __restrict__ out_bwd, long long bwd_stride, int h, int w,
                                                      float clampv) {
    using C = IterCfg<HALO>;
    constexpr int WIN = 2 * HALO + 1;
    extern __shared__ __align__(16) float smem[];
    // layout: s[c][r][col], pitch C::PITCH
    const int pd = blockIdx.z;  // 2*pair + direction
    const int pair = pd >> 1, dir = pd & 1;
    const long long plane = (long long)h * w;
    const float* Rp = R + (long long)(2 * pair) * 5 * plane;
    const float* Rn = Rp + 5 * plane;
    const float* R0 = dir ? Rn : Rp;
    const float* R1 = dir ? Rp : Rn;
    const float2* fin = reinterpret_cast<const float2*>(flow_in) + (long long)pd * plane;
    float2* fout = reinterpret_cast<float2*>(dir ? out_bwd + (long long)pair * bwd_stride
                                                 : out_fwd + (long long)pair * fwd_stride);
    const int x0 = blockIdx.x * IT_TW, y0 = blockIdx.y * IT_TH;
    const int tid = threadIdx.x;
    constexpr int CH = C::RH * C::PITCH;

    // phase 1: M over the halo region (coordinates clamped = replicate border of the box filter)
    for (int idx = tid; idx < C::RH * C::RW; idx += 256) {
        const int r = idx / C::RW, c = idx - r * C::RW;
        const int gy = min(max(y0 + r - HALO, 0), h - 1);
        const int gx = min(max(x0 + c - HALO, 0), w - 1);
        const float2 f = fin[(long long)gy * w + gx];
        float m[5];
        update_matrix_px(R0, R1, plane, w, h, gx, gy, f.x, f.y, m);
        const int so = r * C::PITCH + c;
#pragma unroll
        for (int k = 0; k < 5; ++k) smem[k * CH + so] = m[k];
    }
    __syncthreads();

    // phase 2: vertical window sums, in place: V[r] = sum_{j<WIN} M[r + j], r in [0, TH)
    for (int task = tid; task < 5 * C::RW; task += 256) {
        const int k = task / C::RW, c = task - k * C::RW;
        float* col = smem + k * CH + c;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < WIN; ++j) acc += col[j * C::PITCH];
        float oldest = col[0];
        col[0] = acc;
        for (int r = 1; r < IT_TH; ++r) {
            const float incoming = col[(r + WIN - 1) * C::PITCH];
            acc += incoming - oldest;
            oldest = col[r * C::PITCH];
            col[r * C::PITCH] = acc;
        }
    }
    __syncthreads();

    // phase 3: horizontal window sums for 8 consecutive outputs, solve, store
    {
        const int r = tid >> 3, seg = tid & 7;
        const int gy = y0 + r;
        const int cx = seg * 8;  // first output column within the tile
        float g[5][8];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float* row = smem + k * CH + r * C::PITCH + cx;
            float v[8 + 2 * HALO];
#pragma unroll
            for (int j = 0; j < (8 + 2 * HALO) / 4; ++j) {
                const float4 t = *reinterpret_cast<const float4*>(row + 4 * j);
                v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
            }
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < WIN; ++j) acc += v[j];
            g[k][0] = acc;
#pragma unroll
            for (int i = 1; i < 8; ++i) {
                acc += v[i + WIN - 1] - v[i - 1];
                g[k][i] = acc;
            }
        }
        if (gy < h) {
            const float scale = 1.f / (float)(WIN * WIN);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int gx = x0 + cx + i;
                if (gx < w) {
                    const float g11 = g[0][i] * scale, g12 = g[1][i] * scale, g22 = g[2][i] * scale;
                    const float h1 = g[3][i] * scale, h2 = g[4][i] * scale;
                    const float idet = 1.f / (diff_of_products(g11, g22, g12, g12) + 1e-3f);
                    float fx = diff_of_products(g11, h2, g12, h1) * idet;
                    float fy = diff_of_products(g22, h1, g12, h2) * idet;
                    if (clampv > 0.f) {
                        fx = fminf(fmaxf(fx, -clampv), clampv);
                        fy = fminf(fmaxf(fy, -clampv), clampv);
                    }
                    fout[(long long)gy * w + gx] = make_float2(fx, fy);
                }
            }
        }
    }
}

int launch_fb_iteration(const float* R, const float* flow_in, float* out_fwd, long long fwd_stride, float* out_bwd,
                        long long bwd_stride, int n_pairs, int h, int w, int win, float clamp, bool full_res, cudaStream_t s) {
    if (win != 13) {
        set_error("fb iteration: only winSize 13 is built (got %d)", win);
        return TF_ERR_UNSUPPORTED;
    }
    using C = IterCfg<6>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(fb_iter_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
        attr_set = true;
    }
    const int nz_total = 2 * n_pairs;
    LaunchTimer lt(full_res ? KC_FB_ITER_L0 : KC_FB_ITER, 56.0 * h * w * nz_total, s, cdiv(nz_total, 65534));
    for (int z0 = 0; z0 < nz_total; z0 += 65534) {
        const int nz = min(nz_total - z0, 65534);
        const int p0 = z0 / 2;
        dim3 g(cdiv(w, IT_TW), cdiv(h, IT_TH), nz);
        fb_iter_kernel<6><<<g, 256, C::SMEM_BYTES, s>>>(R + (long long)z0 * 5 * h * w, flow_in + (long long)z0 * 2 * h * w,
                                                        out_fwd + p0 * fwd_stride, fwd_stride, out_bwd + p0 * bwd_stride,
                                                        bwd_stride, h, w, clamp);
    }
    return check_launch("fb iteration");
}

}  // namespace tf
