// K10: variational refinement of a dense flow field (Brox-style), replacing cv2.VariationalRefinement.calc as the
// reference calls it at tobac_flow/flow.py:359,513-519 (vr_model.calc(prev, next, flow) when vr_steps > 0).
//
// OpenCV algorithm (modules/video/src/variational_refinement.cpp, restated and pinned in oracle/varref_np.py):
//   warp I1 by the flow (bilinear, 1/32-px quantised coordinates, replicate border), average with I0, central
//   differences of the averaged image and of the temporal difference -> brightness- and gradient-constancy data terms
//   with robust (Charbonnier) weights, a robust first-order smoothness term whose weights come from the current flow,
//   and red-black SOR on the 2x2-block linear system; 5 fixed-point iterations x 5 SOR iterations by default.
//
// GPU mapping (HBM-bound stencils): the eight derivative images are never stored - every fixed-point iteration
// recomputes them from the averaged image and the temporal difference (8 B/px instead of 32 B/px of reads); the
// linear-system coefficients and the smoothness weights are written once per fixed-point iteration and read by the ten
// SOR half-sweeps.  One thread per pixel (per coloured pixel in the SOR sweeps), grid.z = pair x direction.
#include "tf_common.cuh"

namespace tf {

struct VrArgs {
    const uint8_t* q0; const uint8_t* q1;      // (n_pairs, H, W) quantised frames
    float* fwd; long long fwd_stride;          // flow refined in place, pair p at fwd + p*fwd_stride
    float* bwd; long long bwd_stride;
    float* ws;                                 // workspace: 10 planes per (pair, direction)
    int H, W;
    float alpha2, delta2, gamma2, omega, zeta2, eps2;
};

constexpr int VR_PLANES = 12;   // avg, Iz, A11, A12, A22, b1, b2, sw, (du, dv) x 2 (ping-pong for the fused sweeps)
enum { VP_AVG = 0, VP_IZ, VP_A11, VP_A12, VP_A22, VP_B1, VP_B2, VP_SW, VP_DU, VP_DV, VP_DU2, VP_DV2 };

__device__ __forceinline__ float* vr_plane(const VrArgs& a, int z, int which) {
    return a.ws + ((long long)z * VR_PLANES + which) * ((long long)a.H * a.W);
}
__device__ __forceinline__ const float2* vr_flow(const VrArgs& a, int z) {
    const int p = z >> 1;
    return reinterpret_cast<const float2*>((z & 1) ? a.bwd + p * a.bwd_stride : a.fwd + p * a.fwd_stride);
}

// warp + average + temporal difference; also clears the increment (du, dv)
__global__ void __launch_bounds__(256) vr_prepare_kernel(VrArgs a) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, z = blockIdx.z;
    if (x >= a.W || y >= a.H) return;
    const int H = a.H, W = a.W, p = z >> 1, dir = z & 1;
    const long long hw = (long long)H * W;
    const uint8_t* I0 = (dir ? a.q1 : a.q0) + p * hw;
    const uint8_t* I1 = (dir ? a.q0 : a.q1) + p * hw;
    const int o = y * W + x;
    const float2 f = vr_flow(a, z)[o];
    // remap(I1 as CV_32F, x + u, y + v, INTER_LINEAR, BORDER_REPLICATE)
    const float px = __fadd_rn((float)x, f.x), py = __fadd_rn((float)y, f.y);
    int sx = (fabsf(px) < 6.0e7f) ? __float2int_rn(__fmul_rn(px, 32.f)) : INT_MIN;
    int sy = (fabsf(py) < 6.0e7f) ? __float2int_rn(__fmul_rn(py, 32.f)) : INT_MIN;
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const float fx = (float)(sx & 31) * (1.f / 32.f), fy = (float)(sy & 31) * (1.f / 32.f);
    const int x0 = min(max(ix, 0), W - 1), x1 = min(max(ix + 1, 0), W - 1);
    const int y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
    const float v00 = (float)I1[y0 * W + x0], v01 = (float)I1[y0 * W + x1];
    const float v10 = (float)I1[y1 * W + x0], v11 = (float)I1[y1 * W + x1];
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    const float warped = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, w00), __fmul_rn(v01, w01)), __fmul_rn(v10, w10)),
                                   __fmul_rn(v11, w11));
    const float i0 = (float)I0[o];
    vr_plane(a, z, VP_AVG)[o] = __fadd_rn(__fmul_rn(0.5f, i0), __fmul_rn(0.5f, warped));
    vr_plane(a, z, VP_IZ)[o] = __fsub_rn(warped, i0);
    vr_plane(a, z, VP_DU)[o] = 0.f;
    vr_plane(a, z, VP_DV)[o] = 0.f;
}

// one fixed-point iteration's linear system: data terms + smoothness term
__global__ void __launch_bounds__(256) vr_system_kernel(VrArgs a, int cur) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, z = blockIdx.z;
    if (x >= a.W || y >= a.H) return;
    const int H = a.H, W = a.W;
    const int o = y * W + x;
    const float* __restrict__ avg = vr_plane(a, z, VP_AVG);
    const float* __restrict__ Izp = vr_plane(a, z, VP_IZ);
    const float* __restrict__ dup = vr_plane(a, z, VP_DU + 2 * cur);
    const float* __restrict__ dvp = vr_plane(a, z, VP_DV + 2 * cur);
    const float2* __restrict__ Wf = vr_flow(a, z);
    auto cx = [W](int v) { return min(max(v, 0), W - 1); };
    auto cy = [H](int v) { return min(max(v, 0), H - 1); };
    // Sobel(ksize = 1, BORDER_REPLICATE) first derivatives of the averaged image at arbitrary (clamped) positions
    auto Ixf = [&](int xx, int yy) { return avg[yy * W + cx(xx + 1)] - avg[yy * W + cx(xx - 1)]; };
    auto Iyf = [&](int xx, int yy) { return avg[cy(yy + 1) * W + xx] - avg[cy(yy - 1) * W + xx]; };
    const float Ix = Ixf(x, y), Iy = Iyf(x, y), Iz = Izp[o];
    const float Ixx = Ixf(cx(x + 1), y) - Ixf(cx(x - 1), y);
    const float Ixy = Ixf(x, cy(y + 1)) - Ixf(x, cy(y - 1));
    const float Iyy = Iyf(x, cy(y + 1)) - Iyf(x, cy(y - 1));
    const float Ixz = Izp[y * W + cx(x + 1)] - Izp[y * W + cx(x - 1)];
    const float Iyz = Izp[cy(y + 1) * W + x] - Izp[cy(y - 1) * W + x];
    const float du = dup[o], dv = dvp[o];

    // brightness constancy
    const float rdn = 1.f / (Ix * Ix + Iy * Iy + a.zeta2);
    const float Ik1z = Iz + Ix * du + Iy * dv;
    float wgt = (a.delta2 * rsqrtf(Ik1z * Ik1z * rdn + a.eps2)) * rdn;
    float A11 = wgt * (Ix * Ix) + a.zeta2, A12 = wgt * (Ix * Iy), A22 = wgt * (Iy * Iy) + a.zeta2;
    float b1 = -wgt * (Iz * Ix), b2 = -wgt * (Iz * Iy);
    // gradient constancy
    const float r1 = 1.f / (Ixx * Ixx + Ixy * Ixy + a.zeta2), r2 = 1.f / (Iyy * Iyy + Ixy * Ixy + a.zeta2);
    const float Ik1zx = Ixz + Ixx * du + Ixy * dv, Ik1zy = Iyz + Ixy * du + Iyy * dv;
    wgt = a.gamma2 * rsqrtf(Ik1zx * Ik1zx * r1 + Ik1zy * Ik1zy * r2 + a.eps2);
    A11 += wgt * (Ixx * Ixx * r1 + Ixy * Ixy * r2);
    A12 += wgt * (Ixx * Ixy * r1 + Ixy * Iyy * r2);
    A22 += wgt * (Ixy * Ixy * r1 + Iyy * Iyy * r2);
    b1 -= wgt * (Ixx * Ixz * r1 + Ixy * Iyz * r2);
    b2 -= wgt * (Ixy * Ixz * r1 + Iyy * Iyz * r2);

    // smoothness: weights from the current flow W + dW (forward differences, zero across the image edge),
    // right-hand side from the input flow W
    auto cur_flow = [&](int xx, int yy) {
        const float2 w = Wf[yy * W + xx];
        return make_float2(w.x + dup[yy * W + xx], w.y + dvp[yy * W + xx]);
    };
    auto weight_at = [&](int xx, int yy) {
        const float2 c = cur_flow(xx, yy);
        float ux = 0.f, vx = 0.f, uy = 0.f, vy = 0.f;
        if (xx < W - 1) { const float2 r = cur_flow(xx + 1, yy); ux = r.x - c.x; vx = r.y - c.y; }
        if (yy < H - 1) { const float2 d = cur_flow(xx, yy + 1); uy = d.x - c.x; vy = d.y - c.y; }
        return a.alpha2 * rsqrtf(ux * ux + vx * vx + uy * uy + vy * vy + a.eps2);
    };
    const float sw = weight_at(x, y);
    const float2 w0 = Wf[o];
    const float wxr = (x < W - 1) ? sw : 0.f;                       // edge (x, x+1)
    const float wxl = (x > 0) ? weight_at(x - 1, y) : 0.f;           // edge (x-1, x)
    const float wyd = (y < H - 1) ? sw : 0.f;                       // edge (y, y+1)
    const float wyu = (y > 0) ? weight_at(x, y - 1) : 0.f;           // edge (y-1, y)
    float gxr_u = 0.f, gxr_v = 0.f, gxl_u = 0.f, gxl_v = 0.f, gyd_u = 0.f, gyd_v = 0.f, gyu_u = 0.f, gyu_v = 0.f;
    if (x < W - 1) { const float2 n = Wf[o + 1]; gxr_u = n.x - w0.x; gxr_v = n.y - w0.y; }
    if (x > 0) { const float2 n = Wf[o - 1]; gxl_u = w0.x - n.x; gxl_v = w0.y - n.y; }
    if (y < H - 1) { const float2 n = Wf[o + W]; gyd_u = n.x - w0.x; gyd_v = n.y - w0.y; }
    if (y > 0) { const float2 n = Wf[o - W]; gyu_u = w0.x - n.x; gyu_v = w0.y - n.y; }
    A11 = (((A11 + wxr) + wxl) + wyd) + wyu;
    A22 = (((A22 + wxr) + wxl) + wyd) + wyu;
    b1 = (((b1 + wxr * gxr_u) - wxl * gxl_u) + wyd * gyd_u) - wyu * gyu_u;
    b2 = (((b2 + wxr * gxr_v) - wxl * gxl_v) + wyd * gyd_v) - wyu * gyu_v;

    vr_plane(a, z, VP_A11)[o] = A11;
    vr_plane(a, z, VP_A12)[o] = A12;
    vr_plane(a, z, VP_A22)[o] = A22;
    vr_plane(a, z, VP_B1)[o] = b1;
    vr_plane(a, z, VP_B2)[o] = b2;
    vr_plane(a, z, VP_SW)[o] = sw;
}

// one red or black half-sweep of SOR (in place: the four neighbours of a pixel have the other colour)
__global__ void __launch_bounds__(256) vr_sor_kernel(VrArgs a, int colour, int cur) {
    const int xi = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, z = blockIdx.z;
    const int x = 2 * xi + ((y + colour) & 1);
    if (x >= a.W || y >= a.H) return;
    const int H = a.H, W = a.W;
    const int o = y * W + x;
    float* __restrict__ dup = vr_plane(a, z, VP_DU + 2 * cur);
    float* __restrict__ dvp = vr_plane(a, z, VP_DV + 2 * cur);
    const float* __restrict__ swp = vr_plane(a, z, VP_SW);
    const float sw = swp[o];
    const float wr = (x < W - 1) ? sw : 0.f, wd = (y < H - 1) ? sw : 0.f;
    const float wl = (x > 0) ? swp[o - 1] : 0.f, wu = (y > 0) ? swp[o - W] : 0.f;
    float su = 0.f, sv = 0.f;
    if (x > 0) { su += wl * dup[o - 1]; sv += wl * dvp[o - 1]; }
    if (x < W - 1) { su += wr * dup[o + 1]; sv += wr * dvp[o + 1]; }
    if (y > 0) { su += wu * dup[o - W]; sv += wu * dvp[o - W]; }
    if (y < H - 1) { su += wd * dup[o + W]; sv += wd * dvp[o + W]; }
    const float A12 = vr_plane(a, z, VP_A12)[o];
    float du = dup[o], dv = dvp[o];
    du += a.omega * ((su + vr_plane(a, z, VP_B1)[o] - dv * A12) / vr_plane(a, z, VP_A11)[o] - du);
    dv += a.omega * ((sv + vr_plane(a, z, VP_B2)[o] - du * A12) / vr_plane(a, z, VP_A22)[o] - dv);
    dup[o] = du;
    dvp[o] = dv;
}

// All SOR sweeps of one fixed-point iteration in one launch.  A CTA owns a 32 x 64 output tile and works on the tile
// plus a halo of 2*sweeps pixels: every half-sweep is exact one ring further in, so after `sweeps` red-black
// iterations the inner tile holds what the global sweeps would have produced.  The increment (du, dv) and the
// smoothness weights live in shared memory; the five system coefficients of the (<= 10) pixels a thread owns stay in
// registers for all sweeps.  Each thread owns horizontally adjacent red/black PAIRS, so a half-sweep never diverges.
// HBM bytes per pixel per fixed-point iteration: ~2.1 x 32 B read + 8 B written, against ten half-sweeps of strided
// coefficient reads when the sweeps are separate launches.
constexpr int VS_TW = 64, VS_TH = 32, VS_MAX_SWEEPS = 5;
constexpr int VS_HALO = 2 * VS_MAX_SWEEPS;
constexpr int VS_RW = VS_TW + 2 * VS_HALO;        // 84 (even: pairs never straddle rows)
constexpr int VS_RH = VS_TH + 2 * VS_HALO;        // 52
constexpr int VS_THREADS = 512;
constexpr int VS_PAIRS = VS_RH * (VS_RW / 2);     // 2184
constexpr int VS_PPT = (VS_PAIRS + VS_THREADS - 1) / VS_THREADS;   // 5 pairs per thread
constexpr int VS_SMEM_BYTES = 3 * VS_RH * VS_RW * (int)sizeof(float);

__global__ void __launch_bounds__(VS_THREADS, 2) vr_sor_fused_kernel(VrArgs a, int sweeps, int cur) {
    extern __shared__ __align__(16) float vs_smem[];
    float2* s_uv = reinterpret_cast<float2*>(vs_smem);   // (du, dv) interleaved
    float* s_sw = vs_smem + 2 * VS_RH * VS_RW;
    const int H = a.H, W = a.W, z = blockIdx.z;
    const int x0 = blockIdx.x * VS_TW - VS_HALO, y0 = blockIdx.y * VS_TH - VS_HALO;   // image coords of region (0, 0)
    // reads increment set `cur`, writes the other one: neighbouring CTAs read each other's tiles as halo
    const float* __restrict__ dup = vr_plane(a, z, VP_DU + 2 * cur);
    const float* __restrict__ dvp = vr_plane(a, z, VP_DV + 2 * cur);
    float* __restrict__ dup_out = vr_plane(a, z, VP_DU + 2 * (cur ^ 1));
    float* __restrict__ dvp_out = vr_plane(a, z, VP_DV + 2 * (cur ^ 1));
    const float* __restrict__ swp = vr_plane(a, z, VP_SW);
    const float* __restrict__ pA11 = vr_plane(a, z, VP_A11);
    const float* __restrict__ pA12 = vr_plane(a, z, VP_A12);
    const float* __restrict__ pA22 = vr_plane(a, z, VP_A22);
    const float* __restrict__ pB1 = vr_plane(a, z, VP_B1);
    const float* __restrict__ pB2 = vr_plane(a, z, VP_B2);
    const int tid = threadIdx.x;

    // stage the increment and the weights of the whole region (zero outside the image)
    for (int i = tid; i < VS_RH * VS_RW; i += VS_THREADS) {
        const int r = i / VS_RW, c = i - r * VS_RW;
        const int gy = y0 + r, gx = x0 + c;
        const bool in = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
        const int o = gy * W + gx;
        s_uv[i] = in ? make_float2(dup[o], dvp[o]) : make_float2(0.f, 0.f);
        // the weight of an edge that leaves the image is zero: folded in here so the sweeps need no image-edge tests
        // s_sw keeps the (x, x+1)/(y, y+1) weight of the pixel; the right/bottom image edge is handled per pixel below
        s_sw[i] = in ? swp[o] : 0.f;
    }
    // the coefficients of this thread's pixels
    float cA11[VS_PPT][2], cA12[VS_PPT][2], cA22[VS_PPT][2], cB1[VS_PPT][2], cB2[VS_PPT][2];
#pragma unroll
    for (int j = 0; j < VS_PPT; ++j) {
        const int pidx = tid + j * VS_THREADS;
        const int r = pidx / (VS_RW / 2), c = 2 * (pidx - r * (VS_RW / 2));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int gy = y0 + r, gx = x0 + c + e;
            const bool in = pidx < VS_PAIRS && (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
            const int o = gy * W + gx;
            // omega / A: one division per pixel here instead of two per pixel per half-sweep
            cA11[j][e] = in ? a.omega / pA11[o] : 1.f;
            cA12[j][e] = in ? pA12[o] : 0.f;
            cA22[j][e] = in ? a.omega / pA22[o] : 1.f;
            cB1[j][e] = in ? pB1[o] : 0.f;
            cB2[j][e] = in ? pB2[o] : 0.f;
        }
    }
    __syncthreads();

    const float om1 = 1.f - a.omega;
    for (int half = 0; half < 2 * sweeps; ++half) {
        const int colour = half & 1;
#pragma unroll
        for (int j = 0; j < VS_PPT; ++j) {
            const int pidx = tid + j * VS_THREADS;
            if (pidx >= VS_PAIRS) continue;
            const int r = pidx / (VS_RW / 2), c0 = 2 * (pidx - r * (VS_RW / 2));
            const int gy = y0 + r;
            // the element of the pair whose image coordinates have (x + y) % 2 == colour
            const int e = ((x0 + c0 + gy) & 1) == colour ? 0 : 1;
            const int c = c0 + e, gx = x0 + c;
            if ((unsigned)gy >= (unsigned)H || (unsigned)gx >= (unsigned)W) continue;
            // after half-sweep k only pixels at least k+1 rings inside the region are still exact: skip the rest
            // (this also keeps every neighbour index inside the region)
            if (r <= half || r >= VS_RH - 1 - half || c <= half || c >= VS_RW - 1 - half) continue;
            const int i = r * VS_RW + c;
            const float sw = s_sw[i];
            // pixels outside the image hold zero weights, so only the pixel's own right/bottom edge needs a test
            const float wr = (gx < W - 1) ? sw : 0.f, wd = (gy < H - 1) ? sw : 0.f;
            const float wl = s_sw[i - 1], wu = s_sw[i - VS_RW];
            const float2 nl = s_uv[i - 1], nr = s_uv[i + 1], nu = s_uv[i - VS_RW], nd = s_uv[i + VS_RW];
            const float su = ((wl * nl.x + wr * nr.x) + wu * nu.x) + wd * nd.x;
            const float sv = ((wl * nl.y + wr * nr.y) + wu * nu.y) + wd * nd.y;
            const float A11 = e ? cA11[j][1] : cA11[j][0], A12 = e ? cA12[j][1] : cA12[j][0];
            const float A22 = e ? cA22[j][1] : cA22[j][0], b1 = e ? cB1[j][1] : cB1[j][0], b2 = e ? cB2[j][1] : cB2[j][0];
            float du = s_uv[i].x, dv = s_uv[i].y;
            // du += omega * ((su + b1 - dv*A12) / A11 - du), with A11 holding omega / A11
            du = fmaf(A11, su + b1 - dv * A12, om1 * du);
            dv = fmaf(A22, sv + b2 - du * A12, om1 * dv);
            s_uv[i] = make_float2(du, dv);
        }
        __syncthreads();
    }
    // write the inner tile back
    for (int i = tid; i < VS_TH * VS_TW; i += VS_THREADS) {
        const int r = i / VS_TW, c = i - r * VS_TW;
        const int gy = y0 + VS_HALO + r, gx = x0 + VS_HALO + c;
        if (gy < H && gx < W) {
            const int si = (r + VS_HALO) * VS_RW + c + VS_HALO;
            dup_out[gy * W + gx] = s_uv[si].x;
            dvp_out[gy * W + gx] = s_uv[si].y;
        }
    }
}

// flow <- flow + (du, dv)
__global__ void __launch_bounds__(256) vr_apply_kernel(VrArgs a, int cur) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, z = blockIdx.z;
    if (x >= a.W || y >= a.H) return;
    const int o = y * a.W + x;
    float2* f = const_cast<float2*>(vr_flow(a, z));
    float2 w = f[o];
    w.x += vr_plane(a, z, VP_DU + 2 * cur)[o];
    w.y += vr_plane(a, z, VP_DV + 2 * cur)[o];
    f[o] = w;
}

}  // namespace tf

using namespace tf;

extern "C" void tf_vr_default_params(tf_vr_params* p) {
    if (!p) return;
    p->alpha = 20.f; p->delta = 5.f; p->gamma = 10.f; p->omega = 1.6f;
    p->fixed_point_iterations = 5; p->sor_iterations = 5;
    p->zeta = 0.1f; p->epsilon = 0.001f;
}

extern "C" size_t tf_vr_workspace_bytes(int n_pairs, int H, int W) {
    if (n_pairs <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)2 * n_pairs * VR_PLANES * (size_t)H * W * sizeof(float);
}

extern "C" int tf_variational_refinement(const uint8_t* q0, const uint8_t* q1, float* fwd, long long fwd_stride, float* bwd,
                                         long long bwd_stride, int n_pairs, int H, int W, const tf_vr_params* p,
                                         void* workspace, size_t workspace_bytes, void* stream) {
    if (n_pairs == 0) return TF_OK;
    if (!q0 || !q1 || !fwd || !bwd || !p || !workspace || n_pairs < 0 || H <= 0 || W <= 0) {
        set_error("tf_variational_refinement: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    if ((long long)H * W > 0x3fffffffLL) { set_error("tf_variational_refinement: frame too large"); return TF_ERR_INVALID_ARGUMENT; }
    if (p->fixed_point_iterations < 0 || p->sor_iterations < 0) { set_error("tf_variational_refinement: negative iteration count"); return TF_ERR_INVALID_ARGUMENT; }
    const size_t need = tf_vr_workspace_bytes(n_pairs, H, W);
    if (workspace_bytes < need) {
        set_error("tf_variational_refinement: workspace too small (%zu < %zu)", workspace_bytes, need);
        return TF_ERR_WORKSPACE_TOO_SMALL;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // the attribute is per device (a process may drive several): set it on every call, it is cheap
    cudaFuncSetAttribute(vr_sor_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VS_SMEM_BYTES);
    VrArgs a{};
    a.q0 = q0; a.q1 = q1; a.fwd = fwd; a.fwd_stride = fwd_stride; a.bwd = bwd; a.bwd_stride = bwd_stride;
    a.ws = reinterpret_cast<float*>(workspace); a.H = H; a.W = W;
    a.alpha2 = p->alpha / 2; a.delta2 = p->delta / 2; a.gamma2 = p->gamma / 2; a.omega = p->omega;
    a.zeta2 = p->zeta * p->zeta; a.eps2 = p->epsilon * p->epsilon;
    const double N = (double)H * W * 2 * n_pairs;
    const int launches = 2 + p->fixed_point_iterations * (1 + 2 * p->sor_iterations);
    // algorithmic bytes: prepare 2+8+16, per fixed-point iteration: system 32 read + 24 written, ten half-sweeps of
    // (coefficients 20 + weights 4 + increment 8 read, 8 written) on half the pixels, apply 24
    LaunchTimer lt(KC_VR, N * (26.0 + p->fixed_point_iterations * (56.0 + 2.0 * p->sor_iterations * 20.0) + 24.0), s, launches);
    dim3 block(32, 8);
    for (int z0 = 0; z0 < 2 * n_pairs; z0 += 65534) {
        const int nz = min(2 * n_pairs - z0, 65534);
        VrArgs b = a;
        const int p0 = z0 / 2;
        b.q0 = q0 + (long long)p0 * H * W; b.q1 = q1 + (long long)p0 * H * W;
        b.fwd = fwd + p0 * fwd_stride; b.bwd = bwd + p0 * bwd_stride;
        b.ws = a.ws + (long long)z0 * VR_PLANES * H * W;
        dim3 grid(cdiv(W, 32), cdiv(H, 8), nz), grid_half(cdiv(cdiv(W, 2), 32), cdiv(H, 8), nz);
        dim3 grid_fused(cdiv(W, VS_TW), cdiv(H, VS_TH), nz);
        vr_prepare_kernel<<<grid, block, 0, s>>>(b);
        int cur = 0;   // which (du, dv) set holds the current increment
        for (int fp = 0; fp < p->fixed_point_iterations; ++fp) {
            vr_system_kernel<<<grid, block, 0, s>>>(b, cur);
            if (p->sor_iterations >= 1 && p->sor_iterations <= VS_MAX_SWEEPS) {
                vr_sor_fused_kernel<<<grid_fused, VS_THREADS, VS_SMEM_BYTES, s>>>(b, p->sor_iterations, cur);
                cur ^= 1;
            } else {
                for (int it = 0; it < p->sor_iterations; ++it) {
                    vr_sor_kernel<<<grid_half, block, 0, s>>>(b, 0, cur);
                    vr_sor_kernel<<<grid_half, block, 0, s>>>(b, 1, cur);
                }
            }
        }
        vr_apply_kernel<<<grid, block, 0, s>>>(b, cur);
    }
    return check_launch("tf_variational_refinement");
}
