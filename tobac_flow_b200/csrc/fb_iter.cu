// K4+K5: one fused Farneback iteration = FarnebackUpdateMatrices + FarnebackUpdateFlow_Blur (box window).
//
// OpenCV's sweep is a pure Jacobi update (the matrices it refreshes behind the sliding window are never re-read in the
// same sweep), so flow buffers ping-pong and M is never materialised in HBM.
//
// Strip-marching design (the shape of OpenCV's own sliding window, mapped to a CTA):
//   * a CTA owns a strip of NT columns (NT - 12 outputs + a 6-column halo on each side) and marches down a chunk of
//     rows; there is no row halo re-read apart from a 12-row warm-up per chunk;
//   * M phase, one thread per column, software-pipelined one row ahead (the next row's gathers and the flow of the
//     row after that are in flight while the current row is computed): flow (8 B), R0 (float4 + float) and the
//     bilinear R1 gather at p + flow (4 x (float4 + float)) -> the five M terms.  Rows are handled in batches of 4;
//     the vertical 13-row window sum is assembled from fresh partial sums (prefix sums of batch b-3 kept in a small
//     shared-memory ring and subtracted from its full sum, the full sums of batches b-3..b-1 in registers, the
//     running prefix of batch b), so rounding never accumulates down the chunk;
//   * H phase, every 4 rows: the vertical sums of 4 rows are exchanged through shared memory; each thread takes one row
//     and 4 adjacent outputs, slides the horizontal 13-column window over them, solves the 2x2 systems in registers and
//     stores the new flow (32 B per thread, coalesced).
//   * forward and backward CTAs of the same pair and strip are adjacent in launch order, so the second reader of the
//     shared R planes hits L2.
//   * work distribution: a grid of (strip, row chunk, pair) CTAs, the chunk count chosen to minimise waves x rows.
//     TF_PERSIST=1 selects the measured alternative: all (pair, strip) columns laid end to end and one resident wave of
//     forward/backward CTA pairs taking equal spans of that row space.  It saves warm-up rows and the tail (+2 % with
//     equal code), but its column loop costs the row loop 28 bytes of spills and the spill-free chunk grid is 4 %
//     faster (276 vs 289 ms per CONUS day at the full-resolution level), so the chunk grid is the default.
// Measured and rejected: loading R0 about two rows ahead in place of one (equal: 91.5 vs 91.1 ms); pulling the strip's
// next rows into L2 ahead of the march, either with five
// cp.async.bulk.prefetch.L2 per row or with one prefetch.global.L2 per 128-byte line from warp 0, 4-16 rows ahead
// (117-134 ms vs 95 ms: the extra work of one warp delays the whole CTA at the batch barrier); two rows of taps in flight per thread (168 registers, 3 CTAs/SM: 126-144 ms vs 94.5 ms; a warp
// has six scoreboards, already taken by cur / next taps, the flow queue and the shared-memory reads); inheriting a row's top taps from the previous row's bottom taps (per-lane predicated loads,
// 168 registers, 3 CTAs/SM): 113 ms vs 94.5 ms per 96-frame step for the full-resolution level.
// Algorithmic HBM bytes per pixel-iteration: flow 8 + R0 20 + R1 20 read, flow 8 written = 56 B.
#include <stdlib.h>

#include "farneback_internal.cuh"

namespace tf {

constexpr int IT_HALO = 6, IT_WIN = 13, IT_RB = 4;
#ifndef TF_L2_PREFETCH_ROWS
#define TF_L2_PREFETCH_ROWS 0
#endif
constexpr int IT_PREFETCH_ROWS = TF_L2_PREFETCH_ROWS;
#ifndef TF_DEEP
#define TF_DEEP 0    // 1: a row's tap registers are re-loaded for the row after next as soon as they are blended
#endif
#define TF_PF (TF_DEEP ? 2 : 1)      // rows of taps in flight ahead of the row being computed
#ifndef TF_FQ
#define TF_FQ 4      // rows the flow loads run ahead of the tap issue that consumes them (4 or 2)
#endif
#ifndef TF_HS_CTAS
#define TF_HS_CTAS 4
#endif   // measured: no gain on B200, so off

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// HK = outputs per thread in the H phase = rows per H phase (4: after every M batch, vertical sums double-buffered;
// 8: after every second M batch, single buffer, one more barrier).  Same shared-memory footprint either way.
template <int NT, int HK, bool HS = false, bool TM = false>
struct StripCfg {
    static constexpr int OUT_W = NT - 2 * IT_HALO;
    static constexpr int VPAD = HS ? 0 : 8;                  // the shuffle H phase never reads beyond the strip
    static constexpr int VP = NT + 2 * VPAD;                 // pitch of a row of vertical sums (zero pads both sides)
    // 3 batches x prefix sums P0..P2 x 5 channels; TM: the ring lives in tensor memory (thread-private columns)
    static constexpr int RING_FLOATS = TM ? 0 : 3 * 3 * 5 * NT;
    static constexpr int TM_COLS = 64;                       // 45 used; allocations are powers of two >= 32
    static constexpr int VBUF_ROWS = 8;                      // 2 x 4 (double-buffered) or 1 x 8
    static constexpr int VBUF_FLOATS = VBUF_ROWS * 5 * VP;
    static constexpr int SMEM_BYTES = (RING_FLOATS + VBUF_FLOATS) * (int)sizeof(float);
};

__device__ __forceinline__ float border_factor(int p, int n) {
    // border[] = {0.14, 0.14, 0.4472, 0.4472, 0.4472} applied from both sides
    float s = 1.f;
    if (p < 5) s *= (p < 2 ? 0.14f : 0.4472f);
    const int q = n - 1 - p;
    if (q < 5) s *= (q < 2 ? 0.14f : 0.4472f);
    return s;
}

// Tensor memory as thread-private scratch: with the 32x32b shape lane i of warp w addresses TMEM lane 32 * (w % 4) + i,
// so a column is one private 32-bit word per thread.  The prefix-sum ring of the vertical window lives there: its
// loads and stores then use the TMEM datapath (LDTM / STTM) instead of shared-memory wavefronts of the L1 data pipe,
// which is the unit that limits this kernel.
__device__ __forceinline__ void tm_ld5(float v[5], uint32_t a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(a) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v[4]) : "r"(a + 4) : "memory");
}
__device__ __forceinline__ void tm_st5(uint32_t a, const float v[5]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a + 4), "f"(v[4]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Everything one pixel's FarnebackUpdateMatrices reads: R0 at the pixel, the four bilinear taps of R1 at p + flow.
struct Taps {
    float4 c;  float c4;                 // R0: (c0..c3), c4
    float4 p00, p01, p10, p11;           // R1 float4 plane taps
    float q00, q01, q10, q11;            // R1 c4 plane taps
    float fx, fy, dx, dy;                // bilinear fractions and the flow
    int y;                               // image row (replicate-clamped)
    bool inside;
};

struct RPlanes {
    const float4* R0a; const float* R0b; const float4* R1a; const float* R1b;
};

// issue the loads of one pixel (addresses are always valid; `inside` says whether the R1 taps are used)
__device__ __forceinline__ void issue_taps(Taps& t, const RPlanes& R, int w, int h, int x, int y, float2 f) {
    const int o = y * w + x;
    t.dx = f.x;
    t.dy = f.y;
    float fx = (float)x + f.x, fy = (float)y + f.y;
    const float flx = floorf(fx), fly = floorf(fy);
    const int x1 = (int)flx, y1 = (int)fly;
#if !TF_DEEP
    t.y = y;
    t.fx = fx - flx;
    t.fy = fy - fly;
    t.inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
#endif
    // out-of-image positions read a clamped (valid) 2x2 footprint whose values are then ignored; levels are never
    // narrower than 2 px in practice, and max(.., 0) keeps the address valid even then
#ifndef TF_ABL
#define TF_ABL 0     // timing ablations (wrong results): 1 = taps at the undisplaced pixel, 2 = no R1 loads
#endif
#if TF_ABL >= 1 && TF_ABL <= 3
    const int xc = max(min(x, w - 2), 0), yc = max(min(y, h - 2), 0);
#else
    const int xc = max(min(x1, w - 2), 0), yc = max(min(y1, h - 2), 0);
#endif
    const float4* a0 = R.R1a + (yc * w + xc);
    const float4* a1 = a0 + w;
    const float* b0 = R.R1b + (yc * w + xc);
    const float* b1 = b0 + w;
    t.c = ld_stream(R.R0a + o);      // touched once by this CTA: do not displace the gather footprint in L1
    t.c4 = ld_stream1(R.R0b + o);
#if TF_ABL == 2
    t.p00 = t.p01 = t.p10 = t.p11 = t.c;
    t.q00 = t.q01 = t.q10 = t.q11 = t.c4;
#else
    t.p00 = __ldg(a0);
    t.p01 = __ldg(a0 + 1);
    t.p10 = __ldg(a1);
    t.p11 = __ldg(a1 + 1);
    t.q00 = __ldg(b0);
    t.q01 = __ldg(b0 + 1);
    t.q10 = __ldg(b1);
    t.q11 = __ldg(b1 + 1);
#endif
}

// FarnebackUpdateMatrices for one pixel, part 1: everything that reads the loaded taps -> (r2..r6) before the border
// scaling.  After this the tap registers are dead (TF_DEEP re-issues the loads of the row after next into them).
// LEAN: the fractions / inside flag are recomputed from the flow (same arithmetic as issue_taps) instead of being kept
// in registers next to the in-flight taps.
template <bool LEAN = false>
__device__ __forceinline__ void blend_taps(const Taps& t, float r[5], int w = 0, int h = 0, int x = 0, int y = 0) {
    float r2, r3, r4, r5, r6;
    float fx = t.fx, fy = t.fy;
    bool inside = t.inside;
    if (LEAN) {
        fx = (float)x + t.dx; fy = (float)y + t.dy;
        const float flx = floorf(fx), fly = floorf(fy);
        inside = (unsigned)(int)flx < (unsigned)(w - 1) && (unsigned)(int)fly < (unsigned)(h - 1);
        fx -= flx; fy -= fly;
    }
    if (inside) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * t.p00.x + a01 * t.p01.x + a10 * t.p10.x + a11 * t.p11.x;
        r3 = a00 * t.p00.y + a01 * t.p01.y + a10 * t.p10.y + a11 * t.p11.y;
        r4 = a00 * t.p00.z + a01 * t.p01.z + a10 * t.p10.z + a11 * t.p11.z;
        r5 = a00 * t.p00.w + a01 * t.p01.w + a10 * t.p10.w + a11 * t.p11.w;
        r6 = a00 * t.q00 + a01 * t.q01 + a10 * t.q10 + a11 * t.q11;
        r4 = (t.c.z + r4) * 0.5f;
        r5 = (t.c.w + r5) * 0.5f;
        r6 = (t.c4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = t.c.z;
        r5 = t.c.w;
        r6 = t.c4 * 0.5f;
    }
    r2 = (t.c.x - r2) * 0.5f;
    r3 = (t.c.y - r3) * 0.5f;
    r2 += r4 * t.dy + r6 * t.dx;
    r3 += r6 * t.dy + r5 * t.dx;
    r[0] = r2; r[1] = r3; r[2] = r4; r[3] = r5; r[4] = r6;
}

// part 2: border scaling and the five products
__device__ __forceinline__ void matrix_from_blend(const float r[5], int y, int h, float sc_x, float m[5]) {
    float r2 = r[0], r3 = r[1], r4 = r[2], r5 = r[3], r6 = r[4];
    if (sc_x != 1.f || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float sc = sc_x * border_factor(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

__device__ __forceinline__ void matrix_from_taps(const Taps& t, int h, float sc_x, float m[5]) {
    float r[5];
    blend_taps(t, r);
    matrix_from_blend(r, t.y, h, sc_x, m);
}

// a*b - c*d with one rounding error in the result (Kahan)
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    const float cd = c * d;
    const float err = fmaf(-c, d, cd);
    const float dop = fmaf(a, b, -cd);
    return dop + err;
}

// PS (persistent) selects the work distribution at compile time: the chunk-grid instantiation carries none of the
// column loop's state (it needs the 128-register budget to itself; the persistent one spills 28 bytes).
template <int NT, int HK, bool HS, bool PS, bool TM>
__global__ void __launch_bounds__(NT, (NT == 256 ? 2 : (HS ? TF_HS_CTAS : 4)))
fb_iter_strip_kernel(const float* __restrict__ R, long long img_stride, const float* __restrict__ flow_in,
                     float* __restrict__ out_fwd, long long fwd_stride, float* __restrict__ out_bwd,
                     long long bwd_stride, int h, int w, int chunk_rows, float clampv, int strips, int span_arg, int total) {
    const int span = PS ? span_arg : 0;
    using C = StripCfg<NT, HK, HS, TM>;
    extern __shared__ __align__(16) float smem[];
    float* ring = smem;                       // [batch % 3][P0..P2][k][col]: prefix sums of the batch's rows
    float* vbuf = smem + C::RING_FLOATS;      // [buf][row][k][VP]
    const int tid = threadIdx.x;
    uint32_t tm_base = 0;                     // TM: this warp's lane quarter, column 0 of the CTA's allocation
    if constexpr (TM) {
        static_assert(!TM || NT == 128, "tensor-memory ring: one TMEM lane per thread, 4 warps");
        __shared__ uint32_t tm_addr_s;
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"((uint32_t)__cvta_generic_to_shared(&tm_addr_s)), "n"(C::TM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tm_base = tm_addr_s + ((uint32_t)(tid >> 5) << 21);      // lane (bits 31:16) = 32 * warp
    }
    const int dir = blockIdx.x & 1;
    const int plane = h * w;
    // Work distribution.  span == 0: one (strip, chunk, pair) per CTA from the grid.  span > 0 (persistent): the
    // (pair, strip) columns are laid end to end into one row space of `total` = n_pairs * strips * h rows and CTA pair
    // s (forward + backward) marches rows [s * span, (s + 1) * span) of it, restarting the window at every column
    // boundary it crosses: one wave of equally loaded CTAs, warm-up rows paid once or twice per CTA instead of per chunk.
    // (the span bounds are recomputed from %ctaid at every column instead of being kept live across the row loop)
    auto cta_pair = []() { unsigned v; asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(v)); return (int)(v >> 1); };
    int col = span > 0 ? (cta_pair() * span) / h : 0;
    for (bool first = true;; first = false, ++col) {
    int strip, pair, yc0, yc1;
    if (span > 0) {
        const int lin0 = cta_pair() * span, lin1 = min(lin0 + span, total);
        if (col * h >= lin1) break;
        pair = col / strips;
        strip = col - pair * strips;
        yc0 = max(lin0 - col * h, 0);
        yc1 = min(lin1 - col * h, h);
        if (!first) __syncthreads();          // the previous segment's last H phase is done with the shared buffers
    } else {
        if (!first) break;
        strip = blockIdx.x >> 1;
        pair = blockIdx.z;
        yc0 = blockIdx.y * chunk_rows;
        yc1 = min(yc0 + chunk_rows, h);
    }
    const float* Rp = R + (long long)(2 * pair) * img_stride;
    const float* Rn = Rp + img_stride;
    const float* R0 = dir ? Rn : Rp;
    const float* R1 = dir ? Rp : Rn;
    RPlanes RP;
    RP.R0a = reinterpret_cast<const float4*>(R0);
    RP.R0b = R0 + 4 * (long long)plane;
    RP.R1a = reinterpret_cast<const float4*>(R1);
    RP.R1b = R1 + 4 * (long long)plane;
    const float2* fin = reinterpret_cast<const float2*>(flow_in) + (long long)(2 * pair + dir) * plane;
    float2* fout = reinterpret_cast<float2*>(dir ? out_bwd + (long long)pair * bwd_stride
                                                 : out_fwd + (long long)pair * fwd_stride);
    const int x0 = strip * C::OUT_W;

    // M-phase identity: one column of the strip (replicate-clamped = the box filter's border rule)
    const int gx = min(max(x0 - IT_HALO + tid, 0), w - 1);
    const float sc_x = border_factor(gx, w);
    // H-phase identity: one row of the H batch, HK adjacent output columns
    const int hr = tid / (NT / HK), cg = tid % (NT / HK);

    // zero the suffix-sum ring column and the pads of the vertical-sum rows
    if constexpr (TM) {
        const float z[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 9; ++s) tm_st5(tm_base + 5 * s, z);
        tm_wait_st();
    } else {
#pragma unroll
        for (int s = 0; s < 3 * 3 * 5; ++s) ring[s * NT + tid] = 0.f;
    }
    if constexpr (C::VPAD > 0) {
        for (int i = tid; i < C::VBUF_ROWS * 5 * 2 * C::VPAD; i += NT) {
            const int rowk = i / (2 * C::VPAD), j = i % (2 * C::VPAD);
            vbuf[rowk * C::VP + (j < C::VPAD ? j : NT + j)] = 0.f;
        }
    }
    // the 13-row window of row r0+j is: rows j..3 of batch b-3 (its full sum minus its prefix P_{j-1}, kept in the
    // ring) + batches b-2, b-1 + prefix P_j of batch b.  Every partial sum is formed fresh from at most 4 values, so
    // rounding never accumulates down the chunk.
    float B1[5], B2[5], B3[5];   // full sums of batches b-1, b-2, b-3
#pragma unroll
    for (int k = 0; k < 5; ++k) B1[k] = B2[k] = B3[k] = 0.f;
    // batches of 4 rows are aligned to absolute image rows, so the partial sums a window is built from (and hence
    // the result bits) do not depend on where the chunk starts, i.e. on the launch geometry / batch size
    const int r_begin = (((yc0 - IT_HALO + 8) >> 2) << 2) - 8;
    const int n_rows = (yc1 + IT_HALO) - r_begin;
    const int n_batches = (n_rows + IT_RB - 1) / IT_RB;
    int rb = 0;                  // ring slot of batch b (b % 3): overwritten at the end of the batch, read as b-3 first

    // software pipeline: the current row's taps are in registers, the next row's taps and four rows of flow in flight
    auto row_y = [&](int i) { return min(max(r_begin + i, 0), h - 1); };
    Taps cur;
    issue_taps(cur, RP, w, h, gx, row_y(0), ld_stream(fin + row_y(0) * w + gx));
#if TF_DEEP
    Taps cur1;                   // row i + 1 (the two sets alternate: S[i & 1] holds row i)
    issue_taps(cur1, RP, w, h, gx, row_y(1), ld_stream(fin + row_y(1) * w + gx));
#endif
    float2 fq[TF_FQ];            // flows of rows i+PF .. i+PF+FQ-1 (a DRAM round trip ahead of their use)
#pragma unroll
    for (int j = 0; j < TF_FQ; ++j) fq[j] = ld_stream(fin + row_y(TF_PF + j) * w + gx);
    __syncthreads();

    for (int b = 0; b < n_batches; ++b) {
        // rows of the vertical-sum buffer this M batch fills
        float* vb = vbuf + (HK == 4 ? (b & 1) : 0) * (IT_RB * 5 * C::VP);
        float* vbm = vbuf + (b & 1) * (IT_RB * 5 * C::VP);
        float* rg = ring + rb * (3 * 5 * NT) + tid;
        float P[5], pold[5];
        // ---- M phase: 4 rows of this thread's column -------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < IT_RB; ++j) {
            const int i = b * IT_RB + j;
            // prefetch: the next row's taps (its flow was requested four rows ago) and the flow of row i+5
#if TF_DEEP
            // blend this row's taps, then immediately re-issue the same registers for row i + 2: two rows of gathers in
            // flight per thread with two register sets (the loads of row i + 1 were issued one row ago)
            Taps& st = (j & 1) ? cur1 : cur;
            float rbl[5];
            const int y_cur = row_y(i);
            blend_taps<true>(st, rbl, w, h, gx, y_cur);
            issue_taps(st, RP, w, h, gx, row_y(i + TF_PF), fq[j % TF_FQ]);
            fq[j % TF_FQ] = ld_stream(fin + row_y(i + TF_PF + TF_FQ) * w + gx);
#else
            Taps nxt;
            issue_taps(nxt, RP, w, h, gx, row_y(i + TF_PF), fq[j % TF_FQ]);
            fq[j % TF_FQ] = ld_stream(fin + row_y(i + TF_PF + TF_FQ) * w + gx);
#endif
            if (IT_PREFETCH_ROWS > 0) {
                // pull the R rows this column will gather from a few rows later into L2 (both images: R0 and R1)
                const int op = row_y(i + IT_PREFETCH_ROWS) * w + gx;
                prefetch_l2(RP.R1a + op);
                prefetch_l2(RP.R0a + op);
                if ((tid & 3) == 0) { prefetch_l2(RP.R1b + op); prefetch_l2(RP.R0b + op); }
            }
            float m[5];
#if TF_ABL == 4
            m[0] = cur.p00.x + cur.p01.y + cur.p10.z + cur.p11.w + cur.c.x;
            m[1] = cur.q00 + cur.q01 + cur.q10 + cur.q11 + cur.c4;
            m[2] = cur.dx; m[3] = cur.dy; m[4] = cur.c.w;
#elif TF_DEEP
            matrix_from_blend(rbl, y_cur, h, sc_x, m);
#else
            matrix_from_taps(cur, h, sc_x, m);
#endif
            if constexpr (TM) {
                // same arithmetic, the ring slot in tensor memory: consume P_{j-1} of batch b-3, fetch its P_j (async),
                // write the vertical sums, then overwrite the slot with this batch's P_j
                float xold[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    P[k] = (j == 0) ? m[k] : P[k] + m[k];
                    xold[k] = B3[k];
                    if (j > 0) xold[k] -= pold[k];
                }
                const uint32_t slot = tm_base + (uint32_t)((rb * 3 + j) * 5);
                if (j < IT_RB - 1) tm_ld5(pold, slot);
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    vbm[(j * 5 + k) * C::VP + C::VPAD + tid] = (xold[k] + B2[k]) + (B1[k] + P[k]);
                if (j < IT_RB - 1) {
                    tm_wait_ld();
                    tm_st5(slot, P);
                }
            } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                P[k] = (j == 0) ? m[k] : P[k] + m[k];
                // rows j..3 of batch b-3 = its full sum minus its prefix P_{j-1}; the slot is then reused for batch b
                float xold = B3[k];
                if (j > 0) xold -= pold[k];
                if (j < IT_RB - 1) {
                    pold[k] = rg[(j * 5 + k) * NT];   // P_j of batch b-3, needed by the next row
                    rg[(j * 5 + k) * NT] = P[k];
                }
                vbm[(j * 5 + k) * C::VP + C::VPAD + tid] = (xold + B2[k]) + (B1[k] + P[k]);
            }
            }
#if !TF_DEEP
            cur = nxt;
#endif
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            B3[k] = B2[k];
            B2[k] = B1[k];
            B1[k] = P[k];
        }
        rb = (rb == 2) ? 0 : rb + 1;
        if constexpr (TM) tm_wait_st();
        if (HK == 8 && (b & 1) == 0 && b + 1 < n_batches) continue;   // H phase after every second batch
        __syncthreads();
        // ---- H phase: row hr of the H batch, outputs HK*cg .. HK*cg+HK-1 ---------------------------------------
        const int hb0 = (HK == 8) ? (b & ~1) : b;                      // first M batch covered by this H phase
        const int y = r_begin + hb0 * IT_RB + hr - IT_HALO;
        if (y >= yc0 && y < yc1 && (HK == 4 || hr < IT_RB * (b - hb0 + 1))) {
            float g[5][HK];
            if (TF_ABL == 3 || TF_ABL == 4) {
#pragma unroll
                for (int k = 0; k < 5; ++k)
#pragma unroll
                    for (int i = 0; i < HK; ++i) g[k][i] = 0.f;
            } else
            if constexpr (HS) {
                // a warp owns the whole row (NT / HK == 32): every lane reads only its own quad of vertical sums and the
                // window halves come from the neighbouring lanes as prefix / suffix / full quad sums
                static_assert(!HS || (NT / HK == 32 && HK == 4), "shuffle H phase: one warp per row of quads");
                constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float4 q = *reinterpret_cast<const float4*>(vb + (hr * 5 + k) * C::VP + C::VPAD + HK * cg);
                    const float p2 = q.x + q.y, s2 = q.z + q.w;
                    const float T = p2 + s2, p3 = p2 + q.z, s3 = q.y + s2;
                    const float Tm1 = __shfl_up_sync(FULL, T, 1), Tp1 = __shfl_down_sync(FULL, T, 1);
                    const float s2m2 = __shfl_up_sync(FULL, s2, 2), s1m2 = __shfl_up_sync(FULL, q.w, 2);
                    const float s3m1 = __shfl_up_sync(FULL, s3, 1), p3p1 = __shfl_down_sync(FULL, p3, 1);
                    const float p1p2 = __shfl_down_sync(FULL, q.x, 2), p2p2 = __shfl_down_sync(FULL, p2, 2);
                    const float U = Tm1 + T;
                    g[k][0] = (s2m2 + U) + p3p1;        // columns 4cg-6 .. 4cg+6
                    g[k][1] = (s1m2 + U) + Tp1;         //         4cg-5 .. 4cg+7
                    g[k][2] = (U + Tp1) + p1p2;         //         4cg-4 .. 4cg+8
                    g[k][3] = (s3m1 + T) + (Tp1 + p2p2);  //       4cg-3 .. 4cg+9
                }
            } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                // v[j] = vertical sum at region column HK*cg - 8 + j
                const float4* row = reinterpret_cast<const float4*>(vb + (hr * 5 + k) * C::VP + HK * cg);
                float v[HK + 16];
#pragma unroll
                for (int q = 0; q < (HK + 16) / 4; ++q) {
                    const float4 t = row[q];
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                float acc = 0.f;
#pragma unroll
                for (int j = 2; j < 2 + IT_WIN; ++j) acc += v[j];
                g[k][0] = acc;
#pragma unroll
                for (int i = 1; i < HK; ++i) {
                    acc += v[i + 1 + IT_WIN] - v[i + 1];
                    g[k][i] = acc;
                }
            }
            }
            // OpenCV scales the five sums by 1/169 before the solve; numerator and determinant are both quadratic in
            // them, so the scale folds into the regulariser: 1e-3 * 169^2
            const float reg = 1e-3f * (float)(IT_WIN * IT_WIN) * (float)(IT_WIN * IT_WIN);
            float2 o[HK];
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const float g11 = g[0][i], g12 = g[1][i], g22 = g[2][i];
                const float h1 = g[3][i], h2 = g[4][i];
                const float idet = 1.f / (diff_of_products(g11, g22, g12, g12) + reg);
                float fx = diff_of_products(g11, h2, g12, h1) * idet;
                float fy = diff_of_products(g22, h1, g12, h2) * idet;
                if (clampv > 0.f) {
                    fx = fminf(fmaxf(fx, -clampv), clampv);
                    fy = fminf(fmaxf(fy, -clampv), clampv);
                }
                o[i] = make_float2(fx, fy);
                if (TF_ABL == 4) o[i] = make_float2(1.3f + 1e-30f * fx, 0.7f);
            }
            const int c0 = HK * cg;                        // region column of o[0]
            const int xg = x0 - IT_HALO + c0;              // image column of o[0]
            float2* dst = fout + (long long)y * w + xg;
            // (pairing the outputs into 16-byte stores on even widths measured 2 % slower than four 8-byte stores)
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const int c = c0 + i;
                if (c >= IT_HALO && c < NT - IT_HALO && xg + i < w) dst[i] = o[i];
            }
        }
        if (HK == 8) __syncthreads();   // single vertical-sum buffer: the next M batch overwrites it
    }
    }   // segments
    if constexpr (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_base), "n"(C::TM_COLS) : "memory");
    }
}

template <int NT, int HK, bool HS = false, bool TM = false>
static void launch_strip(const float* R, long long img_stride, const float* flow_in, float* out_fwd, long long fwd_stride,
                         float* out_bwd, long long bwd_stride, int n_pairs, int h, int w, float clamp, cudaStream_t s) {
    using C = StripCfg<NT, HK, HS, TM>;
    // (the attribute is per device: set it on every call, it is cheap)
    cudaFuncSetAttribute(fb_iter_strip_kernel<NT, HK, HS, true, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    cudaFuncSetAttribute(fb_iter_strip_kernel<NT, HK, HS, false, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    const int strips = cdiv(w, C::OUT_W);
    const long long slots = 148LL * (NT == 256 ? 2 : (HS ? TF_HS_CTAS : 4));
    static const char* env_persist = getenv("TF_PERSIST");
    const bool persist = env_persist ? atoi(env_persist) != 0 : false;
    const long long total = (long long)n_pairs * strips * h;
    static const char* env_min = getenv("TF_PERSIST_MIN_ROWS");
    const long long min_rows = env_min ? atoll(env_min) : 128;   // below this the chunk grid's extra parallelism wins
    if (persist && total < 0x40000000LL && total >= min_rows * (slots / 2)) {
        // one resident wave: `slots / 2` forward/backward CTA pairs share the linearised rows equally; at tiny levels a
        // CTA still gets at least 24 rows so the warm-up rows do not dominate
        const int span = (int)max((total + slots / 2 - 1) / (slots / 2), 24LL);
        const int n_cta_pairs = (int)((total + span - 1) / span);
        fb_iter_strip_kernel<NT, HK, HS, true, TM><<<2 * n_cta_pairs, NT, C::SMEM_BYTES, s>>>(
            R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, h, w, 0, clamp, strips, span, (int)total);
        return;
    }
    // rows per chunk: minimise (waves of resident CTAs) x (rows a CTA marches, incl. its 12 warm-up rows)
    int chunks = 1;
    long long best = -1;
    for (int c = 1; c <= max(1, h / 16); ++c) {
        const long long ctas = 2LL * strips * c * n_pairs;
        const long long cost = ((ctas + slots - 1) / slots) * (cdiv(h, c) + 2 * IT_HALO);
        if (best < 0 || cost < best) { best = cost; chunks = c; }
    }
    const int chunk_rows = cdiv(h, chunks);
    chunks = cdiv(h, chunk_rows);
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int np = min(n_pairs - p0, 65535);
        dim3 g(2 * strips, chunks, np);
        fb_iter_strip_kernel<NT, HK, HS, false, TM><<<g, NT, C::SMEM_BYTES, s>>>(R + (long long)(2 * p0) * img_stride, img_stride,
                                                              flow_in + (long long)(2 * p0) * 2 * h * w,
                                                              out_fwd + p0 * fwd_stride, fwd_stride,
                                                              out_bwd + p0 * bwd_stride, bwd_stride, h, w, chunk_rows, clamp,
                                                              strips, 0, 0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// TMA-staged variant (the default): same strip march, same arithmetic (bit-identical results), but
//   * the address-regular streams of a strip -- R0 (float4 + float) and the flow, 28 of the 48 bytes a pixel reads --
//     are staged a batch (4 rows) ahead into shared memory by bulk async copies (cp.async.bulk, TMA engine, one
//     mbarrier per ring slot).  They no longer occupy LSU issue slots, L1 miss-queue entries, scoreboards or
//     destination registers (the flow queue and the R0 halves of both tap sets are gone), and their lookahead is a
//     whole batch instead of one row;
//   * the prefix-sum ring of the vertical window lives in tensor memory (LDTM / STTM instead of shared-memory
//     wavefronts), which is what frees the shared memory for the staging ring at 4 CTAs per SM;
//   * the R1 bilinear gather stays on the L1-cached LDG path (its addresses depend on the flow), PF rows ahead.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct TsCfg {
    static constexpr int NT = 128, HK = 4, OUT_W = NT - 2 * IT_HALO, SLOTS = 2;
    static constexpr int A_ROW = NT * 16;                    // R0 float4 plane: bytes per staged row
    static constexpr int B_ROW = (NT + 4) * 4;               // R0 c4 plane (+ up to 3 floats of alignment skew)
    static constexpr int F_ROW = (NT + 2) * 8;               // flow (+ 1 float2 of alignment skew)
    static constexpr int A_OFF = 0, B_OFF = IT_RB * A_ROW, F_OFF = B_OFF + IT_RB * B_ROW;
    static constexpr int SLOT_BYTES = F_OFF + IT_RB * F_ROW;
    static constexpr int VBUF_FLOATS = 8 * 5 * NT;
    static constexpr int VBUF_OFF = SLOTS * SLOT_BYTES;
    static constexpr int MBAR_OFF = VBUF_OFF + VBUF_FLOATS * 4;
    static constexpr int SMEM_BYTES = MBAR_OFF + SLOTS * 8;
    static constexpr int TM_COLS = 64;
    static_assert(A_ROW % 16 == 0 && B_ROW % 16 == 0 && F_ROW % 16 == 0 && SLOT_BYTES % 16 == 0, "bulk copies: 16-byte granules");
};

// the R1 half of a pixel's reads (R0 and the flow come from the staging ring)
struct TapsR1 {
    float4 p00, p01, p10, p11;
    float q00, q01, q10, q11;
    float dx, dy;
};

__device__ __forceinline__ void issue_taps_r1(TapsR1& t, const float4* __restrict__ R1a, const float* __restrict__ R1b,
                                              int w, int h, int x, int y, float2 f) {
    t.dx = f.x;
    t.dy = f.y;
    const float fx = (float)x + f.x, fy = (float)y + f.y;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    const int xc = max(min(x1, w - 2), 0), yc = max(min(y1, h - 2), 0);
    const float4* a0 = R1a + (yc * w + xc);
    const float4* a1 = a0 + w;
    const float* b0 = R1b + (yc * w + xc);
    const float* b1 = b0 + w;
    t.p00 = __ldg(a0);
    t.p01 = __ldg(a0 + 1);
    t.p10 = __ldg(a1);
    t.p11 = __ldg(a1 + 1);
    t.q00 = __ldg(b0);
    t.q01 = __ldg(b0 + 1);
    t.q10 = __ldg(b1);
    t.q11 = __ldg(b1 + 1);
}

// FarnebackUpdateMatrices part 1 (same arithmetic as blend_taps): taps + R0 (c, c4) -> r2..r6 before border scaling
__device__ __forceinline__ void blend_r1(const TapsR1& t, float4 c, float c4, int w, int h, int x, int y, float r[5]) {
    float fx = (float)x + t.dx, fy = (float)y + t.dy;
    const float flx = floorf(fx), fly = floorf(fy);
    const bool inside = (unsigned)(int)flx < (unsigned)(w - 1) && (unsigned)(int)fly < (unsigned)(h - 1);
    fx -= flx; fy -= fly;
    float r2, r3, r4, r5, r6;
    if (inside) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * t.p00.x + a01 * t.p01.x + a10 * t.p10.x + a11 * t.p11.x;
        r3 = a00 * t.p00.y + a01 * t.p01.y + a10 * t.p10.y + a11 * t.p11.y;
        r4 = a00 * t.p00.z + a01 * t.p01.z + a10 * t.p10.z + a11 * t.p11.z;
        r5 = a00 * t.p00.w + a01 * t.p01.w + a10 * t.p10.w + a11 * t.p11.w;
        r6 = a00 * t.q00 + a01 * t.q01 + a10 * t.q10 + a11 * t.q11;
        r4 = (c.z + r4) * 0.5f;
        r5 = (c.w + r5) * 0.5f;
        r6 = (c4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = c.z;
        r5 = c.w;
        r6 = c4 * 0.5f;
    }
    r2 = (c.x - r2) * 0.5f;
    r3 = (c.y - r3) * 0.5f;
    r2 += r4 * t.dy + r6 * t.dx;
    r3 += r6 * t.dy + r5 * t.dx;
    r[0] = r2; r[1] = r3; r[2] = r4; r[3] = r5; r[4] = r6;
}

template <int PF>
__global__ void __launch_bounds__(TsCfg::NT, 4)
fb_iter_tma_kernel(const float* __restrict__ R, long long img_stride, const float* __restrict__ flow_in,
                   float* __restrict__ out_fwd, long long fwd_stride, float* __restrict__ out_bwd, long long bwd_stride,
                   int h, int w, int chunk_rows, float clampv) {
    using C = TsCfg;
    constexpr int NT = C::NT, HK = C::HK;
    static_assert(PF == 1 || PF == 2, "rows of R1 taps in flight");
    extern __shared__ __align__(128) unsigned char ts_smem[];
    float* vbuf = reinterpret_cast<float*>(ts_smem + C::VBUF_OFF);     // [buf][row][k][NT]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ts_smem + C::MBAR_OFF);
    __shared__ uint32_t tm_addr_s;
    const int tid = threadIdx.x;
    const int dir = blockIdx.x & 1, strip = blockIdx.x >> 1, pair = blockIdx.z;
    const int yc0 = blockIdx.y * chunk_rows, yc1 = min(yc0 + chunk_rows, h);
    const int plane = h * w;
    const float* Rp = R + (long long)(2 * pair) * img_stride;
    const float* Rn = Rp + img_stride;
    const float* R0 = dir ? Rn : Rp;
    const float* R1 = dir ? Rp : Rn;
    const float4* R0a = reinterpret_cast<const float4*>(R0);
    const float* R0b = R0 + 4 * (long long)plane;
    const float4* R1a = reinterpret_cast<const float4*>(R1);
    const float* R1b = R1 + 4 * (long long)plane;
    const float2* fin = reinterpret_cast<const float2*>(flow_in) + (long long)(2 * pair + dir) * plane;
    float2* fout = reinterpret_cast<float2*>(dir ? out_bwd + (long long)pair * bwd_stride
                                                 : out_fwd + (long long)pair * fwd_stride);
    const int x0 = strip * C::OUT_W;
    const int xs = x0 - IT_HALO;                                  // image column of region column 0
    const int cx0 = max(xs, 0), cx1 = min(xs + NT, w), ncol = cx1 - cx0;   // staged (in-image) columns
    const int gx = min(max(xs + tid, 0), w - 1);                  // this thread's column, replicate-clamped
    const int gxs = gx - cx0;                                     // its index inside the staged columns
    const float sc_x = border_factor(gx, w);
    const int hr = tid >> 5, cg = tid & 31;
    // alignment skews of the 4-byte / 8-byte streams (bulk copies move 16-byte granules from 16-byte-aligned addresses)
    const unsigned skb0 = (unsigned)(reinterpret_cast<uintptr_t>(R0b) >> 2), skf0 = (unsigned)(reinterpret_cast<uintptr_t>(fin) >> 3);

    // tensor memory for the prefix-sum ring, barriers of the staging ring
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tm_addr_s)), "n"(C::TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_addr_s + ((uint32_t)(tid >> 5) << 21);
    {
        const float z[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 9; ++s) tm_st5(tm_base + 5 * s, z);
        tm_wait_st();
    }
    float B1[5], B2[5], B3[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) B1[k] = B2[k] = B3[k] = 0.f;
    const int r_begin = (((yc0 - IT_HALO + 8) >> 2) << 2) - 8;
    const int n_rows = (yc1 + IT_HALO) - r_begin;
    const int n_batches = (n_rows + IT_RB - 1) / IT_RB;
    int rb = 0;
    auto row_y = [&](int i) { return min(max(r_begin + i, 0), h - 1); };

    // Stage batch bb (rows 4bb .. 4bb+3, replicate-clamped) into ring slot bb & 1: warp r copies row r, its lanes 0..2
    // one stream each (cp.async.bulk takes warp-uniform operands, so the copies of a warp issue one after the other:
    // three per warp keeps the four warps level at the batch barrier).  Byte counts are constants of the strip -- the
    // 4-byte / 8-byte streams start at the 16-byte boundary below their first element and always move the padded row
    // -- so the slot's barrier is armed with a constant by thread 0.
    const int swarp = tid >> 5, slane = tid & 31;
    const uint32_t bytes_a = (uint32_t)ncol * 16u, bytes_b = (((uint32_t)ncol + 6u) & ~3u) * 4u,
                   bytes_f = (((uint32_t)ncol + 2u) & ~1u) * 8u;
    const uint32_t stage_tx = IT_RB * (bytes_a + bytes_b + bytes_f);
    const char* st_base = slane == 0 ? reinterpret_cast<const char*>(R0a)
                        : (slane == 1 ? reinterpret_cast<const char*>(R0b) : reinterpret_cast<const char*>(fin));
    const int st_esz = slane == 0 ? 16 : (slane == 1 ? 4 : 8);
    const uint32_t st_bytes = slane == 0 ? bytes_a : (slane == 1 ? bytes_b : bytes_f);
    const uint32_t st_dst = smem_u32(ts_smem) + (slane == 0 ? C::A_OFF + swarp * C::A_ROW + (cx0 - xs) * 16
                                                 : (slane == 1 ? C::B_OFF + swarp * C::B_ROW : C::F_OFF + swarp * C::F_ROW));
    auto stage_batch = [&](int bb) {
        const uint32_t bar = smem_u32(&mbar[bb & 1]);
        if (slane < 3) {
            const int y = row_y(IT_RB * bb + swarp);
            const uintptr_t src = (reinterpret_cast<uintptr_t>(st_base) + (long long)(y * w + cx0) * st_esz) & ~(uintptr_t)15;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(st_dst + (uint32_t)(bb & 1) * C::SLOT_BYTES), "l"(src), "r"(st_bytes), "r"(bar) : "memory");
        }
        if (tid == 0) mbar_arrive_expect_tx(&mbar[bb & 1], stage_tx);
    };
    // this thread's staged values of row (bb, j)
    auto staged_flow = [&](int bb, int j) {
        const int y = row_y(IT_RB * bb + j);
        const unsigned sk = (skf0 + (unsigned)(y * w + cx0)) & 1u;
        return reinterpret_cast<const float2*>(ts_smem + (bb & 1) * C::SLOT_BYTES + C::F_OFF + j * C::F_ROW)[sk + gxs];
    };

    stage_batch(0);
    if (n_batches > 1) stage_batch(1);
    mbar_wait(&mbar[0], 0u);
    TapsR1 S0, S1;                                               // PF == 1: S0 = current row, S1 = next row
    issue_taps_r1(S0, R1a, R1b, w, h, gx, row_y(0), staged_flow(0, 0));
    if (PF == 2) issue_taps_r1(S1, R1a, R1b, w, h, gx, row_y(1), staged_flow(0, 1));

    for (int b = 0; b < n_batches; ++b) {
        float* vbm = vbuf + (b & 1) * (IT_RB * 5 * NT);
        const unsigned char* slot = ts_smem + (b & 1) * C::SLOT_BYTES;
        float P[5], pold[5];
#pragma unroll
        for (int j = 0; j < IT_RB; ++j) {
            const int i = b * IT_RB + j;
            const int y_cur = row_y(i);
            // R0 of this row from the staging ring
            const float4 c = reinterpret_cast<const float4*>(slot + C::A_OFF + j * C::A_ROW)[gx - xs];
            const unsigned skb = (skb0 + (unsigned)(y_cur * w + cx0)) & 3u;
            const float c4 = reinterpret_cast<const float*>(slot + C::B_OFF + j * C::B_ROW)[skb + gxs];
            // the flow of row i + PF (the next batch's slot from row 4 - PF on: its copies were issued a batch ago)
            const int jn = j + PF;
            const int bn = b + (jn >= IT_RB ? 1 : 0);
            if (jn == IT_RB && bn < n_batches) mbar_wait(&mbar[bn & 1], (uint32_t)((bn >> 1) & 1));
            float2 fnext = make_float2(0.f, 0.f);
            if (bn < n_batches) fnext = staged_flow(bn, jn & (IT_RB - 1));
            float rbl[5];
            if (PF == 1) {
                issue_taps_r1(S1, R1a, R1b, w, h, gx, row_y(i + 1), fnext);
                blend_r1(S0, c, c4, w, h, gx, y_cur, rbl);
            } else {
                TapsR1& st = (j & 1) ? S1 : S0;
                blend_r1(st, c, c4, w, h, gx, y_cur, rbl);
                issue_taps_r1(st, R1a, R1b, w, h, gx, row_y(i + 2), fnext);
            }
            float m[5];
            matrix_from_blend(rbl, y_cur, h, sc_x, m);
            float xold[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                P[k] = (j == 0) ? m[k] : P[k] + m[k];
                xold[k] = B3[k];
                if (j > 0) xold[k] -= pold[k];
            }
            const uint32_t tslot = tm_base + (uint32_t)((rb * 3 + j) * 5);
            if (j < IT_RB - 1) tm_ld5(pold, tslot);
#pragma unroll
            for (int k = 0; k < 5; ++k) vbm[(j * 5 + k) * NT + tid] = (xold[k] + B2[k]) + (B1[k] + P[k]);
            if (j < IT_RB - 1) {
                tm_wait_ld();
                tm_st5(tslot, P);
            }
            if (PF == 1) S0 = S1;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            B3[k] = B2[k];
            B2[k] = B1[k];
            B1[k] = P[k];
        }
        rb = (rb == 2) ? 0 : rb + 1;
        tm_wait_st();
        __syncthreads();
        // every thread is done with slot b & 1: refill it with batch b + 2
        if (b + 2 < n_batches) stage_batch(b + 2);
        // ---- H phase: a warp owns row hr of the batch, a lane four adjacent outputs ---------------------------------
        const int y = r_begin + b * IT_RB + hr - IT_HALO;
        if (y >= yc0 && y < yc1) {
            float g[5][HK];
            constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float4 q = *reinterpret_cast<const float4*>(vbm + (hr * 5 + k) * NT + HK * cg);
                const float p2 = q.x + q.y, s2 = q.z + q.w;
                const float T = p2 + s2, p3 = p2 + q.z, s3 = q.y + s2;
                const float Tm1 = __shfl_up_sync(FULL, T, 1), Tp1 = __shfl_down_sync(FULL, T, 1);
                const float s2m2 = __shfl_up_sync(FULL, s2, 2), s1m2 = __shfl_up_sync(FULL, q.w, 2);
                const float s3m1 = __shfl_up_sync(FULL, s3, 1), p3p1 = __shfl_down_sync(FULL, p3, 1);
                const float p1p2 = __shfl_down_sync(FULL, q.x, 2), p2p2 = __shfl_down_sync(FULL, p2, 2);
                const float U = Tm1 + T;
                g[k][0] = (s2m2 + U) + p3p1;
                g[k][1] = (s1m2 + U) + Tp1;
                g[k][2] = (U + Tp1) + p1p2;
                g[k][3] = (s3m1 + T) + (Tp1 + p2p2);
            }
            const float reg = 1e-3f * (float)(IT_WIN * IT_WIN) * (float)(IT_WIN * IT_WIN);
            float2 o[HK];
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const float g11 = g[0][i], g12 = g[1][i], g22 = g[2][i];
                const float h1 = g[3][i], h2 = g[4][i];
                const float idet = 1.f / (diff_of_products(g11, g22, g12, g12) + reg);
                float fx = diff_of_products(g11, h2, g12, h1) * idet;
                float fy = diff_of_products(g22, h1, g12, h2) * idet;
                if (clampv > 0.f) {
                    fx = fminf(fmaxf(fx, -clampv), clampv);
                    fy = fminf(fmaxf(fy, -clampv), clampv);
                }
                o[i] = make_float2(fx, fy);
            }
            const int c0 = HK * cg;
            const int xg = xs + c0;
            float2* dst = fout + (long long)y * w + xg;
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const int cc = c0 + i;
                if (cc >= IT_HALO && cc < NT - IT_HALO && xg + i < w) dst[i] = o[i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_addr_s), "n"(C::TM_COLS) : "memory");
}

template <int PF, int WAL>
__global__ void fb_iter_v3_kernel(const float* __restrict__ R, long long img_stride, const float* __restrict__ flow_in,
                                  float* __restrict__ out_fwd, long long fwd_stride, float* __restrict__ out_bwd,
                                  long long bwd_stride, int h, int w, int chunk_rows, float clampv);

template <int PF, bool V3>
static void launch_tma(const float* R, long long img_stride, const float* flow_in, float* out_fwd, long long fwd_stride,
                       float* out_bwd, long long bwd_stride, int n_pairs, int h, int w, float clamp, cudaStream_t s) {
    using C = TsCfg;
    auto kern = V3 ? ((w & 3) == 0 ? fb_iter_v3_kernel<PF, 2> : ((w & 1) == 0 ? fb_iter_v3_kernel<PF, 1> : fb_iter_v3_kernel<PF, 0>))
                   : fb_iter_tma_kernel<PF>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    const int strips = cdiv(w, C::OUT_W);
    const long long slots = 148LL * 4;
    // rows per chunk: minimise (waves of resident CTAs) x (rows a CTA marches, incl. its 12 warm-up rows)
    int chunks = 1;
    long long best = -1;
    for (int c = 1; c <= max(1, h / 16); ++c) {
        const long long ctas = 2LL * strips * c * n_pairs;
        const long long cost = ((ctas + slots - 1) / slots) * (cdiv(h, c) + 2 * IT_HALO);
        if (best < 0 || cost < best) { best = cost; chunks = c; }
    }
    const int chunk_rows = cdiv(h, chunks);
    chunks = cdiv(h, chunk_rows);
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int np = min(n_pairs - p0, 65535);
        dim3 g(2 * strips, chunks, np);
        kern<<<g, C::NT, C::SMEM_BYTES, s>>>(R + (long long)(2 * p0) * img_stride, img_stride,
                                                              flow_in + (long long)(2 * p0) * 2 * h * w,
                                                              out_fwd + p0 * fwd_stride, fwd_stride,
                                                              out_bwd + p0 * bwd_stride, bwd_stride, h, w, chunk_rows, clamp);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed-math kernel (the default).  The strip march is instruction-issue bound (IPC 2.3 of 4 at 16 warps per SM, DRAM
// at 36 %: removing the whole R1 gather buys 12 %, removing the H-phase arithmetic 17 %), so this version cuts warp
// instructions rather than bytes:
//   * Blackwell's packed fp32 pipe (FFMA2 / FADD2 / FMUL2, one issue slot for two lanes of IEEE fp32 arithmetic):
//     the bilinear blend works on the (c0, c1) and (c2, c3) halves of every float4 tap, the five normal-equation
//     terms travel as two pairs + one scalar, A = (M0, M2) -> (g11, g22), B = (M3, M4) -> (h1, h2), C = M1 -> g12,
//     through the prefix sums, the vertical window, shared memory (8-byte stores, pairs interleaved per column) and the
//     horizontal window of the H phase.  Every operation is the same IEEE operation on the same operands in the same
//     order as in the scalar kernel above, so the window sums are bit-identical to it;
//   * the prefix-sum ring of the vertical window lives in tensor memory (see tm_ld5 / tm_st5);
//   * the reciprocal of the (regularised, always normal) determinant is one MUFU.RCP (<= 1 ulp) instead of the IEEE
//     division sequence with its slow-path branch.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef TF_RCP_EXACT
#define TF_RCP_EXACT 0
#endif
typedef float2 f2;
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 dup2(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 lo2(float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ f2 hi2(float4 v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ f2 shfl_up2(f2 v, int d) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ f2 shfl_down2(f2 v, int d) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d));
}
__device__ __forceinline__ float rcp_fast(float x) {
#if TF_RCP_EXACT
    return 1.f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

struct PkCfg {
    static constexpr int NT = 128, HK = 4, OUT_W = NT - 2 * IT_HALO;
    static constexpr int ROW_FLOATS = 5 * NT;                 // [A: NT x float2][B: NT x float2][C: NT x float]
    static constexpr int VBUF_FLOATS = 8 * ROW_FLOATS;        // 2 buffers x 4 rows
    static constexpr int SMEM_BYTES = VBUF_FLOATS * 4;
    static constexpr int TM_COLS = 64;
};

// FarnebackUpdateMatrices for one pixel: (A, B, C) = ((M0, M2), (M3, M4), M1)
__device__ __forceinline__ void matrix_pk(const Taps& t, int y, int h, float sc_x, f2& mA, f2& mB, float& mC) {
    f2 r23, r45;
    float r6;
    const f2 half2 = make_float2(0.5f, 0.5f), minus1 = make_float2(-1.f, -1.f);
    if (t.inside) {
        const float fx = t.fx, fy = t.fy;
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const f2 w00 = dup2(a00), w01 = dup2(a01), w10 = dup2(a10), w11 = dup2(a11);
        r23 = fma2(w11, lo2(t.p11), fma2(w10, lo2(t.p10), fma2(w01, lo2(t.p01), mul2(w00, lo2(t.p00)))));
        r45 = fma2(w11, hi2(t.p11), fma2(w10, hi2(t.p10), fma2(w01, hi2(t.p01), mul2(w00, hi2(t.p00)))));
        r6 = a00 * t.q00 + a01 * t.q01 + a10 * t.q10 + a11 * t.q11;
        r45 = mul2(add2(hi2(t.c), r45), half2);
        r6 = (t.c4 + r6) * 0.25f;
    } else {
        r23 = make_float2(0.f, 0.f);
        r45 = hi2(t.c);
        r6 = t.c4 * 0.5f;
    }
    r23 = mul2(fma2(r23, minus1, lo2(t.c)), half2);             // (c - r) * 0.5
    float r2 = r23.x, r3 = r23.y, r4 = r45.x, r5 = r45.y;
    r2 += r4 * t.dy + r6 * t.dx;
    r3 += r6 * t.dy + r5 * t.dx;
    if (sc_x != 1.f || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float sc = sc_x * border_factor(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    mA.x = r4 * r4 + r6 * r6;
    mC = (r4 + r5) * r6;
    mA.y = r5 * r5 + r6 * r6;
    mB.x = r4 * r2 + r6 * r3;
    mB.y = r6 * r2 + r5 * r3;
}

__device__ __forceinline__ void tm_ld_abc(f2& a, f2& b, float& c, uint32_t addr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a.x), "=f"(a.y), "=f"(b.x), "=f"(b.y) : "r"(addr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(c) : "r"(addr + 4) : "memory");
}
__device__ __forceinline__ void tm_st_abc(uint32_t addr, f2 a, f2 b, float c) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(addr + 4), "f"(c) : "memory");
}

__global__ void __launch_bounds__(PkCfg::NT, 4)
fb_iter_pk_kernel(const float* __restrict__ R, long long img_stride, const float* __restrict__ flow_in,
                  float* __restrict__ out_fwd, long long fwd_stride, float* __restrict__ out_bwd, long long bwd_stride,
                  int h, int w, int chunk_rows, float clampv) {
    using C = PkCfg;
    constexpr int NT = C::NT, HK = C::HK;
    extern __shared__ __align__(16) float pk_vbuf[];             // [buf][row][A | B | C]
    __shared__ uint32_t tm_addr_s;
    const int tid = threadIdx.x;
    const int dir = blockIdx.x & 1, strip = blockIdx.x >> 1, pair = blockIdx.z;
    const int yc0 = blockIdx.y * chunk_rows, yc1 = min(yc0 + chunk_rows, h);
    const int plane = h * w;
    const float* Rp = R + (long long)(2 * pair) * img_stride;
    const float* Rn = Rp + img_stride;
    const float* R0 = dir ? Rn : Rp;
    const float* R1 = dir ? Rp : Rn;
    RPlanes RP;
    RP.R0a = reinterpret_cast<const float4*>(R0);
    RP.R0b = R0 + 4 * (long long)plane;
    RP.R1a = reinterpret_cast<const float4*>(R1);
    RP.R1b = R1 + 4 * (long long)plane;
    const float2* fin = reinterpret_cast<const float2*>(flow_in) + (long long)(2 * pair + dir) * plane;
    float2* fout = reinterpret_cast<float2*>(dir ? out_bwd + (long long)pair * bwd_stride
                                                 : out_fwd + (long long)pair * fwd_stride);
    const int x0 = strip * C::OUT_W;
    const int gx = min(max(x0 - IT_HALO + tid, 0), w - 1);
    const float sc_x = border_factor(gx, w);
    const int hr = tid >> 5, cg = tid & 31;

    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"((uint32_t)__cvta_generic_to_shared(&tm_addr_s)), "n"(C::TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_addr_s + ((uint32_t)(tid >> 5) << 21);
    const f2 zero2 = make_float2(0.f, 0.f), minus1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int s = 0; s < 9; ++s) tm_st_abc(tm_base + 5 * s, zero2, zero2, 0.f);
    tm_wait_st();

    // full sums of batches b-1, b-2, b-3
    f2 B1a = zero2, B1b = zero2, B2a = zero2, B2b = zero2, B3a = zero2, B3b = zero2;
    float B1c = 0.f, B2c = 0.f, B3c = 0.f;
    const int r_begin = (((yc0 - IT_HALO + 8) >> 2) << 2) - 8;
    const int n_rows = (yc1 + IT_HALO) - r_begin;
    const int n_batches = (n_rows + IT_RB - 1) / IT_RB;
    int rb = 0;
    auto row_y = [&](int i) { return min(max(r_begin + i, 0), h - 1); };
    Taps cur;
    issue_taps(cur, RP, w, h, gx, row_y(0), ld_stream(fin + row_y(0) * w + gx));
    float2 fq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) fq[j] = ld_stream(fin + row_y(1 + j) * w + gx);

    for (int b = 0; b < n_batches; ++b) {
        float* vbm = pk_vbuf + (b & 1) * (IT_RB * C::ROW_FLOATS);
        f2 Pa, Pb, olda, oldb;
        float Pc, oldc;
#pragma unroll
        for (int j = 0; j < IT_RB; ++j) {
            const int i = b * IT_RB + j;
            Taps nxt;
            issue_taps(nxt, RP, w, h, gx, row_y(i + 1), fq[j]);
            fq[j] = ld_stream(fin + row_y(i + 5) * w + gx);
            f2 mA, mB;
            float mC;
            matrix_pk(cur, cur.y, h, sc_x, mA, mB, mC);
            if (j == 0) { Pa = mA; Pb = mB; Pc = mC; }
            else { Pa = add2(Pa, mA); Pb = add2(Pb, mB); Pc = Pc + mC; }
            // rows j..3 of batch b-3 = its full sum minus its prefix P_{j-1}
            f2 xa = B3a, xb = B3b;
            float xc = B3c;
            if (j > 0) { xa = fma2(olda, minus1, xa); xb = fma2(oldb, minus1, xb); xc -= oldc; }
            const uint32_t tslot = tm_base + (uint32_t)((rb * 3 + j) * 5);
            if (j < IT_RB - 1) tm_ld_abc(olda, oldb, oldc, tslot);
            float* vr = vbm + j * C::ROW_FLOATS;
            reinterpret_cast<f2*>(vr)[tid] = add2(add2(xa, B2a), add2(B1a, Pa));
            reinterpret_cast<f2*>(vr + 2 * NT)[tid] = add2(add2(xb, B2b), add2(B1b, Pb));
            vr[4 * NT + tid] = (xc + B2c) + (B1c + Pc);
            if (j < IT_RB - 1) {
                tm_wait_ld();
                tm_st_abc(tslot, Pa, Pb, Pc);
            }
            cur = nxt;
        }
        B3a = B2a; B3b = B2b; B3c = B2c;
        B2a = B1a; B2b = B1b; B2c = B1c;
        B1a = Pa; B1b = Pb; B1c = Pc;
        rb = (rb == 2) ? 0 : rb + 1;
        tm_wait_st();
        __syncthreads();
        // ---- H phase: a warp owns row hr of the batch, a lane four adjacent outputs ---------------------------------
        const int y = r_begin + b * IT_RB + hr - IT_HALO;
        if (y >= yc0 && y < yc1) {
            const float* vr = vbm + hr * C::ROW_FLOATS;
            f2 gA[HK], gB[HK];
            float gC[HK];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                const float4* src = reinterpret_cast<const float4*>(vr + pr * 2 * NT) + 2 * cg;
                const float4 u = src[0], v = src[1];
                const f2 q0 = lo2(u), q1 = hi2(u), q2 = lo2(v), q3 = hi2(v);
                const f2 p2 = add2(q0, q1), s2 = add2(q2, q3);
                const f2 T = add2(p2, s2), p3 = add2(p2, q2), s3 = add2(q1, s2);
                const f2 Tm1 = shfl_up2(T, 1), Tp1 = shfl_down2(T, 1);
                const f2 s2m2 = shfl_up2(s2, 2), s1m2 = shfl_up2(q3, 2);
                const f2 s3m1 = shfl_up2(s3, 1), p3p1 = shfl_down2(p3, 1);
                const f2 p1p2 = shfl_down2(q0, 2), p2p2 = shfl_down2(p2, 2);
                const f2 U = add2(Tm1, T);
                f2* g = pr ? gB : gA;
                g[0] = add2(add2(s2m2, U), p3p1);
                g[1] = add2(add2(s1m2, U), Tp1);
                g[2] = add2(add2(U, Tp1), p1p2);
                g[3] = add2(add2(s3m1, T), add2(Tp1, p2p2));
            }
            {
                constexpr unsigned FULL = 0xffffffffu;
                const float4 q = reinterpret_cast<const float4*>(vr + 4 * NT)[cg];
                const float p2 = q.x + q.y, s2 = q.z + q.w;
                const float T = p2 + s2, p3 = p2 + q.z, s3 = q.y + s2;
                const float Tm1 = __shfl_up_sync(FULL, T, 1), Tp1 = __shfl_down_sync(FULL, T, 1);
                const float s2m2 = __shfl_up_sync(FULL, s2, 2), s1m2 = __shfl_up_sync(FULL, q.w, 2);
                const float s3m1 = __shfl_up_sync(FULL, s3, 1), p3p1 = __shfl_down_sync(FULL, p3, 1);
                const float p1p2 = __shfl_down_sync(FULL, q.x, 2), p2p2 = __shfl_down_sync(FULL, p2, 2);
                const float U = Tm1 + T;
                gC[0] = (s2m2 + U) + p3p1;
                gC[1] = (s1m2 + U) + Tp1;
                gC[2] = (U + Tp1) + p1p2;
                gC[3] = (s3m1 + T) + (Tp1 + p2p2);
            }
            const float reg = 1e-3f * (float)(IT_WIN * IT_WIN) * (float)(IT_WIN * IT_WIN);
            float2 o[HK];
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const float g11 = gA[i].x, g22 = gA[i].y, g12 = gC[i];
                const float h1 = gB[i].x, h2 = gB[i].y;
                const float idet = rcp_fast(diff_of_products(g11, g22, g12, g12) + reg);
                float fx = diff_of_products(g11, h2, g12, h1) * idet;
                float fy = diff_of_products(g22, h1, g12, h2) * idet;
                if (clampv > 0.f) {
                    fx = fminf(fmaxf(fx, -clampv), clampv);
                    fy = fminf(fmaxf(fy, -clampv), clampv);
                }
                o[i] = make_float2(fx, fy);
            }
            const int c0 = HK * cg;
            const int xg = x0 - IT_HALO + c0;
            float2* dst = fout + (long long)y * w + xg;
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const int cc = c0 + i;
                if (cc >= IT_HALO && cc < NT - IT_HALO && xg + i < w) dst[i] = o[i];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_addr_s), "n"(C::TM_COLS) : "memory");
}

static void launch_pk(const float* R, long long img_stride, const float* flow_in, float* out_fwd, long long fwd_stride,
                      float* out_bwd, long long bwd_stride, int n_pairs, int h, int w, float clamp, cudaStream_t s) {
    using C = PkCfg;
    const int strips = cdiv(w, C::OUT_W);
    const long long slots = 148LL * 4;
    // rows per chunk: minimise (waves of resident CTAs) x (rows a CTA marches, incl. its 12 warm-up rows)
    int chunks = 1;
    long long best = -1;
    for (int c = 1; c <= max(1, h / 16); ++c) {
        const long long ctas = 2LL * strips * c * n_pairs;
        const long long cost = ((ctas + slots - 1) / slots) * (cdiv(h, c) + 2 * IT_HALO);
        if (best < 0 || cost < best) { best = cost; chunks = c; }
    }
    const int chunk_rows = cdiv(h, chunks);
    chunks = cdiv(h, chunk_rows);
    for (int p0 = 0; p0 < n_pairs; p0 += 65535) {
        const int np = min(n_pairs - p0, 65535);
        dim3 g(2 * strips, chunks, np);
        fb_iter_pk_kernel<<<g, C::NT, C::SMEM_BYTES, s>>>(R + (long long)(2 * p0) * img_stride, img_stride,
                                                         flow_in + (long long)(2 * p0) * 2 * h * w,
                                                         out_fwd + p0 * fwd_stride, fwd_stride,
                                                         out_bwd + p0 * bwd_stride, bwd_stride, h, w, chunk_rows, clamp);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// v3 (the default): TMA-staged regular streams + tensor-memory prefix ring (as fb_iter_tma_kernel, which made the march
// issue-bound: IPC 2.9, long-scoreboard 0.6 per issue) + fewer warp instructions:
//   * packed fp32 (FFMA2 / FADD2 / FMUL2) for the bilinear blend and for the five normal-equation terms, which travel as
//     A = (M0, M2) -> (g11, g22), B = (M3, M4) -> (h1, h2), C = M1 -> g12 through prefix sums, vertical window,
//     shared memory (8-byte stores) and the H phase's horizontal window: same IEEE operations in the same order as the
//     scalar kernels, so the window sums are bit-identical to theirs;
//   * a tap set keeps its bilinear fractions and inside flag (the flow itself is re-read from the staging ring when the
//     row is blended) instead of recomputing floor / int conversions;
//   * one MUFU.RCP for the reciprocal of the regularised (always normal) determinant.
// ---------------------------------------------------------------------------------------------------------------------
struct TapsV3 {
    float4 p00, p01, p10, p11;
    float q00, q01, q10, q11;
    float fx, fy;
    bool inside;
};

__device__ __forceinline__ void issue_taps_v3(TapsV3& t, const float4* __restrict__ R1a, const float* __restrict__ R1b,
                                              int w, int h, int x, int y, float2 f) {
    const float px = (float)x + f.x, py = (float)y + f.y;
    const float flx = floorf(px), fly = floorf(py);
    const int x1 = (int)flx, y1 = (int)fly;
    t.fx = px - flx;
    t.fy = py - fly;
    t.inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    // the taps of a position outside the image are never used: predicated loads instead of clamped addresses
    if (t.inside) {
        const float4* a0 = R1a + (y1 * w + x1);
        const float4* a1 = a0 + w;
        const float* b0 = R1b + (y1 * w + x1);
        const float* b1 = b0 + w;
        t.p00 = __ldg(a0);
        t.p01 = __ldg(a0 + 1);
        t.p10 = __ldg(a1);
        t.p11 = __ldg(a1 + 1);
        t.q00 = __ldg(b0);
        t.q01 = __ldg(b0 + 1);
        t.q10 = __ldg(b1);
        t.q11 = __ldg(b1 + 1);
    }
}

// FarnebackUpdateMatrices for one pixel: (A, B, C) = ((M0, M2), (M3, M4), M1); c / c4 = R0 at the pixel, f = its flow
__device__ __forceinline__ void matrix_v3(const TapsV3& t, float4 c, float c4, float2 f, int y, int h, float sc_x, f2& mA,
                                          f2& mB, float& mC) {
    f2 r23, r45;
    float r6;
    const f2 half2 = make_float2(0.5f, 0.5f), minus1 = make_float2(-1.f, -1.f);
    if (t.inside) {
        const float fx = t.fx, fy = t.fy;
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const f2 w00 = dup2(a00), w01 = dup2(a01), w10 = dup2(a10), w11 = dup2(a11);
        r23 = fma2(w11, lo2(t.p11), fma2(w10, lo2(t.p10), fma2(w01, lo2(t.p01), mul2(w00, lo2(t.p00)))));
        r45 = fma2(w11, hi2(t.p11), fma2(w10, hi2(t.p10), fma2(w01, hi2(t.p01), mul2(w00, hi2(t.p00)))));
        r6 = a00 * t.q00 + a01 * t.q01 + a10 * t.q10 + a11 * t.q11;
        r45 = mul2(add2(hi2(c), r45), half2);
        r6 = (c4 + r6) * 0.25f;
    } else {
        r23 = make_float2(0.f, 0.f);
        r45 = hi2(c);
        r6 = c4 * 0.5f;
    }
    r23 = mul2(fma2(r23, minus1, lo2(c)), half2);                // (c - r) * 0.5
    float r2 = r23.x, r3 = r23.y, r4 = r45.x, r5 = r45.y;
    r2 += r4 * f.y + r6 * f.x;
    r3 += r6 * f.y + r5 * f.x;
    if (sc_x != 1.f || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float sc = sc_x * border_factor(y, h);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    mA.x = r4 * r4 + r6 * r6;
    mC = (r4 + r5) * r6;
    mA.y = r5 * r5 + r6 * r6;
    mB.x = r4 * r2 + r6 * r3;
    mB.y = r6 * r2 + r5 * r3;
}

// WAL = 2: w % 4 == 0 -- every row of the 4-byte / 8-byte streams then starts at the same offset from a 16-byte
// boundary, so the alignment skews of the staged rows are constants of the strip (hoisted) and the outputs can go out as
// 16-byte stores.  WAL = 1: w even -- the same for the 8-byte streams (flow in, flow out) only.  WAL = 0: any w.
template <int PF, int WAL>
__global__ void __launch_bounds__(TsCfg::NT, 4)
fb_iter_v3_kernel(const float* __restrict__ R, long long img_stride, const float* __restrict__ flow_in,
                  float* __restrict__ out_fwd, long long fwd_stride, float* __restrict__ out_bwd, long long bwd_stride,
                  int h, int w, int chunk_rows, float clampv) {
    using C = TsCfg;
    constexpr int NT = C::NT, HK = C::HK, ROWF = 5 * NT;         // a row of vertical sums: [A: NT x f2][B: NT x f2][C: NT]
    static_assert(PF == 1 || PF == 2, "rows of R1 taps in flight");
    extern __shared__ __align__(128) unsigned char ts_smem[];
    float* vbuf = reinterpret_cast<float*>(ts_smem + C::VBUF_OFF);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ts_smem + C::MBAR_OFF);
    __shared__ uint32_t tm_addr_s;
    const int tid = threadIdx.x;
    const int dir = blockIdx.x & 1, strip = blockIdx.x >> 1, pair = blockIdx.z;
    const int yc0 = blockIdx.y * chunk_rows, yc1 = min(yc0 + chunk_rows, h);
    const int plane = h * w;
    const float* Rp = R + (long long)(2 * pair) * img_stride;
    const float* Rn = Rp + img_stride;
    const float* R0 = dir ? Rn : Rp;
    const float* R1 = dir ? Rp : Rn;
    const float4* R0a = reinterpret_cast<const float4*>(R0);
    const float* R0b = R0 + 4 * (long long)plane;
    const float4* R1a = reinterpret_cast<const float4*>(R1);
    const float* R1b = R1 + 4 * (long long)plane;
    const float2* fin = reinterpret_cast<const float2*>(flow_in) + (long long)(2 * pair + dir) * plane;
    float2* fout = reinterpret_cast<float2*>(dir ? out_bwd + (long long)pair * bwd_stride
                                                 : out_fwd + (long long)pair * fwd_stride);
    const int x0 = strip * C::OUT_W;
    const int xs = x0 - IT_HALO;
    const int cx0 = max(xs, 0), cx1 = min(xs + NT, w), ncol = cx1 - cx0;
    const int gx = min(max(xs + tid, 0), w - 1);
    const int gxs = gx - cx0;
    const float sc_x = border_factor(gx, w);
    const int hr = tid >> 5, cg = tid & 31;
    const unsigned skb0 = (unsigned)(reinterpret_cast<uintptr_t>(R0b) >> 2), skf0 = (unsigned)(reinterpret_cast<uintptr_t>(fin) >> 3);

    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tm_addr_s)), "n"(C::TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (tid == 0) {
            mbar_init(&mbar[0], 1);
            mbar_init(&mbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm_base = tm_addr_s + ((uint32_t)(tid >> 5) << 21);
    const f2 zero2 = make_float2(0.f, 0.f), minus1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int s = 0; s < 9; ++s) tm_st_abc(tm_base + 5 * s, zero2, zero2, 0.f);
    tm_wait_st();
    f2 B1a = zero2, B1b = zero2, B2a = zero2, B2b = zero2, B3a = zero2, B3b = zero2;
    float B1c = 0.f, B2c = 0.f, B3c = 0.f;
    const int r_begin = (((yc0 - IT_HALO + 8) >> 2) << 2) - 8;
    const int n_rows = (yc1 + IT_HALO) - r_begin;
    const int n_batches = (n_rows + IT_RB - 1) / IT_RB;
    int rb = 0;
    auto row_y = [&](int i) { return min(max(r_begin + i, 0), h - 1); };

    // staging: see fb_iter_tma_kernel
    const int swarp = tid >> 5, slane = tid & 31;
    const uint32_t bytes_a = (uint32_t)ncol * 16u, bytes_b = (((uint32_t)ncol + 6u) & ~3u) * 4u,
                   bytes_f = (((uint32_t)ncol + 2u) & ~1u) * 8u;
    const uint32_t stage_tx = IT_RB * (bytes_a + bytes_b + bytes_f);
    const char* st_base = slane == 0 ? reinterpret_cast<const char*>(R0a)
                        : (slane == 1 ? reinterpret_cast<const char*>(R0b) : reinterpret_cast<const char*>(fin));
    const int st_esz = slane == 0 ? 16 : (slane == 1 ? 4 : 8);
    const uint32_t st_bytes = slane == 0 ? bytes_a : (slane == 1 ? bytes_b : bytes_f);
    const uint32_t st_dst = smem_u32(ts_smem) + (slane == 0 ? C::A_OFF + swarp * C::A_ROW + (cx0 - xs) * 16
                                                 : (slane == 1 ? C::B_OFF + swarp * C::B_ROW : C::F_OFF + swarp * C::F_ROW));
    auto stage_batch = [&](int bb) {
        const uint32_t bar = smem_u32(&mbar[bb & 1]);
        if (slane < 3) {
            const int y = row_y(IT_RB * bb + swarp);
            const uintptr_t src = (reinterpret_cast<uintptr_t>(st_base) + (long long)(y * w + cx0) * st_esz) & ~(uintptr_t)15;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(st_dst + (uint32_t)(bb & 1) * C::SLOT_BYTES), "l"(src), "r"(st_bytes), "r"(bar) : "memory");
        }
        if (tid == 0) mbar_arrive_expect_tx(&mbar[bb & 1], stage_tx);
    };
    const unsigned skb_c = (skb0 + (unsigned)cx0) & 3u, skf_c = (skf0 + (unsigned)cx0) & 1u;   // WA: the skews of every row
    auto staged_flow = [&](int bb, int j) {
        unsigned sk = skf_c;
        if (WAL == 0) sk = (skf0 + (unsigned)(row_y(IT_RB * bb + j) * w + cx0)) & 1u;
        return reinterpret_cast<const float2*>(ts_smem + (bb & 1) * C::SLOT_BYTES + C::F_OFF + j * C::F_ROW)[sk + gxs];
    };

    stage_batch(0);
    if (n_batches > 1) stage_batch(1);
    mbar_wait(&mbar[0], 0u);
    TapsV3 S0, S1;
    issue_taps_v3(S0, R1a, R1b, w, h, gx, row_y(0), staged_flow(0, 0));
    if (PF == 2) issue_taps_v3(S1, R1a, R1b, w, h, gx, row_y(1), staged_flow(0, 1));

    for (int b = 0; b < n_batches; ++b) {
        float* vbm = vbuf + (b & 1) * (IT_RB * ROWF);
        const unsigned char* slot = ts_smem + (b & 1) * C::SLOT_BYTES;
        f2 Pa, Pb, olda, oldb;
        float Pc, oldc;
#pragma unroll
        for (int j = 0; j < IT_RB; ++j) {
            const int i = b * IT_RB + j;
            const int y_cur = row_y(i);
            // R0 and the flow of this row from the staging ring
            const float4 c = reinterpret_cast<const float4*>(slot + C::A_OFF + j * C::A_ROW)[gx - xs];
            unsigned skb = skb_c, skf = skf_c;
            if (WAL < 2) {
                const unsigned e_cur = (unsigned)(y_cur * w + cx0);
                skb = (skb0 + e_cur) & 3u;
                if (WAL == 0) skf = (skf0 + e_cur) & 1u;
            }
            const float c4 = reinterpret_cast<const float*>(slot + C::B_OFF + j * C::B_ROW)[skb + gxs];
            const float2 fcur = reinterpret_cast<const float2*>(slot + C::F_OFF + j * C::F_ROW)[skf + gxs];
            // the flow of row i + PF (from row 4 - PF on it lives in the next batch's slot, staged a batch ago)
            const int jn = j + PF;
            const int bn = b + (jn >= IT_RB ? 1 : 0);
            if (jn == IT_RB && bn < n_batches) mbar_wait(&mbar[bn & 1], (uint32_t)((bn >> 1) & 1));
            float2 fnext = make_float2(0.f, 0.f);
            if (bn < n_batches) fnext = staged_flow(bn, jn & (IT_RB - 1));
            f2 mA, mB;
            float mC;
            if (PF == 1) {
                issue_taps_v3(S1, R1a, R1b, w, h, gx, row_y(i + 1), fnext);
                matrix_v3(S0, c, c4, fcur, y_cur, h, sc_x, mA, mB, mC);
            } else {
                TapsV3& st = (j & 1) ? S1 : S0;
                matrix_v3(st, c, c4, fcur, y_cur, h, sc_x, mA, mB, mC);
                issue_taps_v3(st, R1a, R1b, w, h, gx, row_y(i + 2), fnext);
            }
            if (j == 0) { Pa = mA; Pb = mB; Pc = mC; }
            else { Pa = add2(Pa, mA); Pb = add2(Pb, mB); Pc = Pc + mC; }
            f2 xa = B3a, xb = B3b;
            float xc = B3c;
            if (j > 0) { xa = fma2(olda, minus1, xa); xb = fma2(oldb, minus1, xb); xc -= oldc; }
            const uint32_t tslot = tm_base + (uint32_t)((rb * 3 + j) * 5);
            if (j < IT_RB - 1) tm_ld_abc(olda, oldb, oldc, tslot);
            float* vr = vbm + j * ROWF;
            reinterpret_cast<f2*>(vr)[tid] = add2(add2(xa, B2a), add2(B1a, Pa));
            reinterpret_cast<f2*>(vr + 2 * NT)[tid] = add2(add2(xb, B2b), add2(B1b, Pb));
            vr[4 * NT + tid] = (xc + B2c) + (B1c + Pc);
            if (j < IT_RB - 1) {
                tm_wait_ld();
                tm_st_abc(tslot, Pa, Pb, Pc);
            }
            if (PF == 1) S0 = S1;
        }
        B3a = B2a; B3b = B2b; B3c = B2c;
        B2a = B1a; B2b = B1b; B2c = B1c;
        B1a = Pa; B1b = Pb; B1c = Pc;
        rb = (rb == 2) ? 0 : rb + 1;
        tm_wait_st();
        __syncthreads();
        if (b + 2 < n_batches) stage_batch(b + 2);
        // ---- H phase: a warp owns row hr of the batch, a lane four adjacent outputs ---------------------------------
        const int y = r_begin + b * IT_RB + hr - IT_HALO;
        if (y >= yc0 && y < yc1) {
            const float* vr = vbm + hr * ROWF;
            f2 gA[HK], gB[HK];
            float gC[HK];
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                const float4* src = reinterpret_cast<const float4*>(vr + pr * 2 * NT) + 2 * cg;
                const float4 u = src[0], v = src[1];
                const f2 q0 = lo2(u), q1 = hi2(u), q2 = lo2(v), q3 = hi2(v);
                const f2 p2 = add2(q0, q1), s2 = add2(q2, q3);
                const f2 T = add2(p2, s2), p3 = add2(p2, q2), s3 = add2(q1, s2);
                const f2 Tm1 = shfl_up2(T, 1), Tp1 = shfl_down2(T, 1);
                const f2 s2m2 = shfl_up2(s2, 2), s1m2 = shfl_up2(q3, 2);
                const f2 s3m1 = shfl_up2(s3, 1), p3p1 = shfl_down2(p3, 1);
                const f2 p1p2 = shfl_down2(q0, 2), p2p2 = shfl_down2(p2, 2);
                const f2 U = add2(Tm1, T);
                f2* g = pr ? gB : gA;
                g[0] = add2(add2(s2m2, U), p3p1);
                g[1] = add2(add2(s1m2, U), Tp1);
                g[2] = add2(add2(U, Tp1), p1p2);
                g[3] = add2(add2(s3m1, T), add2(Tp1, p2p2));
            }
            {
                constexpr unsigned FULL = 0xffffffffu;
                const float4 q = reinterpret_cast<const float4*>(vr + 4 * NT)[cg];
                const float p2 = q.x + q.y, s2 = q.z + q.w;
                const float T = p2 + s2, p3 = p2 + q.z, s3 = q.y + s2;
                const float Tm1 = __shfl_up_sync(FULL, T, 1), Tp1 = __shfl_down_sync(FULL, T, 1);
                const float s2m2 = __shfl_up_sync(FULL, s2, 2), s1m2 = __shfl_up_sync(FULL, q.w, 2);
                const float s3m1 = __shfl_up_sync(FULL, s3, 1), p3p1 = __shfl_down_sync(FULL, p3, 1);
                const float p1p2 = __shfl_down_sync(FULL, q.x, 2), p2p2 = __shfl_down_sync(FULL, p2, 2);
                const float U = Tm1 + T;
                gC[0] = (s2m2 + U) + p3p1;
                gC[1] = (s1m2 + U) + Tp1;
                gC[2] = (U + Tp1) + p1p2;
                gC[3] = (s3m1 + T) + (Tp1 + p2p2);
            }
            const float reg = 1e-3f * (float)(IT_WIN * IT_WIN) * (float)(IT_WIN * IT_WIN);
            float2 o[HK];
#pragma unroll
            for (int i = 0; i < HK; ++i) {
                const float g11 = gA[i].x, g22 = gA[i].y, g12 = gC[i];
                const float h1 = gB[i].x, h2 = gB[i].y;
                const float idet = rcp_fast(diff_of_products(g11, g22, g12, g12) + reg);
                float fx = diff_of_products(g11, h2, g12, h1) * idet;
                float fy = diff_of_products(g22, h1, g12, h2) * idet;
                if (clampv > 0.f) {
                    fx = fminf(fmaxf(fx, -clampv), clampv);
                    fy = fminf(fmaxf(fy, -clampv), clampv);
                }
                o[i] = make_float2(fx, fy);
            }
            const int c0 = HK * cg;
            const int xg = xs + c0;
            float2* dst = fout + (long long)y * w + xg;
            if (WAL >= 1 && (reinterpret_cast<uintptr_t>(fout) & 15) == 0) {
                // xg and w are even: the output pairs (0, 1) and (2, 3) are 16-byte aligned and valid / invalid together
#pragma unroll
                for (int i = 0; i < HK; i += 2) {
                    const int cc = c0 + i;
                    if (cc >= IT_HALO && cc < NT - IT_HALO && xg + i < w)
                        *reinterpret_cast<float4*>(dst + i) = make_float4(o[i].x, o[i].y, o[i + 1].x, o[i + 1].y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < HK; ++i) {
                    const int cc = c0 + i;
                    if (cc >= IT_HALO && cc < NT - IT_HALO && xg + i < w) dst[i] = o[i];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm_addr_s), "n"(C::TM_COLS) : "memory");
}

int launch_fb_iteration(const float* R, long long img_stride, const float* flow_in, float* out_fwd, long long fwd_stride,
                        float* out_bwd, long long bwd_stride, int n_pairs, int h, int w, int win, float clamp,
                        bool full_res, cudaStream_t s) {
    if (win != IT_WIN) {
        set_error("fb iteration: only winSize 13 is built (got %d)", win);
        return TF_ERR_UNSUPPORTED;
    }
    if ((long long)h * w > 0x3fffffffLL) { set_error("fb iteration: level too large"); return TF_ERR_INVALID_ARGUMENT; }
    LaunchTimer lt(full_res ? KC_FB_ITER_L0 : KC_FB_ITER, 56.0 * h * w * 2 * n_pairs, s, cdiv(n_pairs, 65535));
    // strip width: the configuration that wastes fewer columns
    const int pad256 = cdiv(w, StripCfg<256, 4>::OUT_W) * 256, pad128 = cdiv(w, StripCfg<128, 4>::OUT_W) * 128;
    static const char* force_hk = getenv("TF_FORCE_HK");
    const bool hk8 = force_hk ? (atoi(force_hk) == 8) : false;
    // 128-column strips (4 resident CTAs per SM) measured faster than 256-column ones (2 per SM) at equal padding:
    // more independent CTAs hide each other's barrier and gather latency.  256 only when it wastes clearly less.
    static const char* force_hs = getenv("TF_HSHFL");
    const bool hshfl = force_hs ? (atoi(force_hs) != 0) : true;
    // TMA-staged kernel: needs 16-byte-aligned float4 planes / flow rows are handled by skews; 2-px-wide levels are not
    static const char* force_tma = getenv("TF_TMA");
    const int tma = force_tma ? atoi(force_tma) : 3;      // 3 = v3 (default); 0 = the scalar strip kernels below
    if (tma && w >= 2 && h >= 2) {
        if (tma == 3) launch_tma<2, true>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
        else if (tma == 4) launch_tma<1, true>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
        else if (tma == 2) launch_tma<2, false>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
        else launch_tma<1, false>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
        return check_launch("fb iteration (tma)");
    }
    static const char* force_pk = getenv("TF_PK");
    if (force_pk && atoi(force_pk) != 0) {
        launch_pk(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
        return check_launch("fb iteration (packed)");
    }
    static const char* force_tm = getenv("TF_TMEM");
    const bool tmem = force_tm ? (atoi(force_tm) != 0) : false;
    static const char* force_nt = getenv("TF_FORCE_NT");
    const bool use256 = force_nt ? (atoi(force_nt) == 256) : (10 * pad256 < 9 * pad128);
    if (use256 && hk8)
        launch_strip<256, 8>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    else if (use256)
        launch_strip<256, 4>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    else if (hk8)
        launch_strip<128, 8>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    else if (hshfl && tmem)
        launch_strip<128, 4, true, true>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    else if (hshfl)
        launch_strip<128, 4, true>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    else
        launch_strip<128, 4>(R, img_stride, flow_in, out_fwd, fwd_stride, out_bwd, bwd_stride, n_pairs, h, w, clamp, s);
    return check_launch("fb iteration");
}

}  // namespace tf
