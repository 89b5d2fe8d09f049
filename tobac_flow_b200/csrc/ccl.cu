// K11: per-frame 2-D connected components (flat_label) and hole filling.
//
// Replaces scipy.ndimage.label as flat_label drives it (tobac_flow/utils/label_utils.py:143-180: the structure's time
// links are removed, so every frame is labelled on its own) and scipy.ndimage.binary_fill_holes with the 2-D cross
// structure (tobac_flow/detection.py:72-87, 330-346).  Label numbers follow scipy exactly: components are numbered in
// raster order (t, y, x) of their first pixel.
//
// Union-find on the pixel grid with run compression: a pixel starts out pointing at the first pixel of its horizontal
// run inside its 32-pixel warp segment (one ballot), so the only unions left are one per segment boundary and one per
// (run, run-above) adjacency.  Roots are always the smallest linear index of their tree (the larger root is hooked
// under the smaller with atomicMin), so after flattening every pixel carries the raster-first pixel of its component
// and numbering is a prefix count of root pixels.
#include "tf_common.cuh"

namespace tf {

__device__ __forceinline__ int uf_find(int* L, int a) {
    volatile int* V = L;
    int p = V[a];
    while (p != a) { a = p; p = V[a]; }
    return a;
}

__device__ __forceinline__ void uf_unite(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// L[p] = first pixel of p's run inside its 32-pixel segment (foreground) or -1 (background); p is frame-local
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W,
                                                       int invert) {
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
    const long long base = (long long)blockIdx.z * H * W;
    const bool in = x < W && y < H;
    const int p = y * W + x;
    const bool fg = in && ((mask[base + p] != 0) != (invert != 0));
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    if (!in) return;
    int v = -1;
    if (fg) {
        const unsigned zeros_below = ~bits & ((1u << lane) - 1u);
        const int start = zeros_below ? 32 - __clz(zeros_below) : 0;
        v = p - (lane - start);
    }
    L[base + p] = v;
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(int* __restrict__ Lall, int H, int W, int conn8) {
    const int lane = threadIdx.x;
    const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
    int* L = Lall + (long long)blockIdx.z * H * W;
    const bool in = x < W && y < H;
    const int p = y * W + x;
    const bool fg = in && L[p] >= 0;   // foreground entries are indices (>= 0) for the whole kernel, background stays -1
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    if (!fg) return;
    const bool left_fg = lane > 0 ? ((bits >> (lane - 1)) & 1u) : (x > 0 && L[p - 1] >= 0);
    if (lane == 0 && left_fg) uf_unite(L, p, p - 1);               // run continues across the segment boundary
    if (y == 0) return;
    const bool seg_start = lane == 0 || !((bits >> (lane - 1)) & 1u);
    const bool up_fg = L[p - W] >= 0;
    const bool upleft_fg = x > 0 && L[p - W - 1] >= 0;
    if (up_fg) {
        if (seg_start || !upleft_fg) uf_unite(L, p, p - W);       // one union per (run, run-above) adjacency
    } else if (conn8) {
        if (upleft_fg && !left_fg) uf_unite(L, p, p - W - 1);
        const bool right_fg = lane < 31 ? ((bits >> (lane + 1)) & 1u) : (x < W - 1 && L[p + 1] >= 0);
        if (x < W - 1 && !right_fg && L[p - W + 1] >= 0) uf_unite(L, p, p - W + 1);
    }
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(int* __restrict__ Lall, long long hw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    int* L = Lall + (long long)blockIdx.y * hw;
    int a = L[i];
    if (a < 0) return;
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    L[i] = a;
}

// ---- scipy numbering: rank of every root pixel in (t, y, x) order --------------------------------------------------
constexpr int NUM_BLK = 1024;

__global__ void __launch_bounds__(NUM_BLK) ccl_count_roots_kernel(const int* __restrict__ Lall, long long hw, int nb,
                                                                  int* __restrict__ blk) {
    const long long i = (long long)blockIdx.x * NUM_BLK + threadIdx.x;
    const int* L = Lall + (long long)blockIdx.y * hw;
    const int is_root = i < hw && L[i] == (int)i;
    const int c = __syncthreads_count(is_root);
    if (threadIdx.x == 0) blk[(long long)blockIdx.y * nb + blockIdx.x] = c;
}

// exclusive scan of `n` ints per row (blockIdx.x = row) in place, row total to tot[row]
__global__ void __launch_bounds__(1024) row_exclusive_scan_kernel(int* __restrict__ v, int n, int* __restrict__ tot) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    int* row = v + (long long)blockIdx.x * n;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int val = i < n ? row[i] : 0;
        int inc = val;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid ? warp_sums[wid - 1] : 0) + inc - val;
        if (i < n) row[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && tot) tot[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(NUM_BLK) ccl_assign_kernel(const int* __restrict__ Lall, long long hw, int nb,
                                                             const int* __restrict__ blk_off, const int* __restrict__ frame_off,
                                                             int* __restrict__ labels) {
    __shared__ int warp_cnt[32];
    const long long i = (long long)blockIdx.x * NUM_BLK + threadIdx.x;
    const int* L = Lall + (long long)blockIdx.y * hw;
    const bool is_root = i < hw && L[i] == (int)i;
    const unsigned b = __ballot_sync(0xffffffffu, is_root);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    if (wid == 0) {
        int c = warp_cnt[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= o) c += t;
        }
        warp_cnt[lane] = c;   // inclusive
    }
    __syncthreads();
    if (is_root) {
        const int rank = (wid ? warp_cnt[wid - 1] : 0) + __popc(b & ((1u << lane) - 1u));
        labels[(long long)blockIdx.y * hw + i] =
            frame_off[blockIdx.y] + blk_off[(long long)blockIdx.y * nb + blockIdx.x] + rank + 1;
    }
}

__global__ void __launch_bounds__(256) ccl_final_kernel(const int* __restrict__ Lall, long long hw, int* __restrict__ labels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const long long base = (long long)blockIdx.y * hw;
    const int r = Lall[base + i];
    if (r < 0) labels[base + i] = 0;
    else if (r != (int)i) labels[base + i] = labels[base + r];   // the root's number was written by the previous kernel
}

__global__ void store_total_kernel(const int* __restrict__ tot, int* __restrict__ dst) { *dst = *tot; }

// ---- hole filling --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ccl_border_flag_kernel(const int* __restrict__ Lall, int H, int W,
                                                              uint8_t* __restrict__ flag) {
    // one thread per border pixel: 2W + 2H per frame
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = 2 * W + 2 * H;
    if (i >= n) return;
    int x, y;
    if (i < W) { x = i; y = 0; }
    else if (i < 2 * W) { x = i - W; y = H - 1; }
    else if (i < 2 * W + H) { x = 0; y = i - 2 * W; }
    else { x = W - 1; y = i - 2 * W - H; }
    const long long base = (long long)blockIdx.y * H * W;
    const int r = Lall[base + y * W + x];
    if (r >= 0) flag[base + r] = 1;
}

__global__ void __launch_bounds__(256) fill_holes_kernel(const int* __restrict__ Lall, const uint8_t* __restrict__ flag,
                                                         long long hw, uint8_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const long long base = (long long)blockIdx.y * hw;
    const int r = Lall[base + i];           // background component (of the inverted mask) or -1 for original foreground
    out[base + i] = (r < 0 || !flag[base + r]) ? 1 : 0;
}

struct CclWs {
    int* L;
    uint8_t* flag;
    int* blk;        // T * nb
    int* frame_tot;  // T
    int* total;      // 1
    size_t bytes;
    int nb;
};

static CclWs carve_ccl(void* base, int T, int H, int W) {
    const size_t N = (size_t)T * H * W;
    const int nb = (int)(((size_t)H * W + NUM_BLK - 1) / NUM_BLK);
    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off += align_up(b, 256); return o; };
    const size_t oL = take(N * 4), oF = take(N), oB = take((size_t)T * nb * 4), oT = take((size_t)T * 4), oS = take(256);
    CclWs w{};
    char* b = reinterpret_cast<char*>(base);
    w.L = reinterpret_cast<int*>(b + oL);
    w.flag = reinterpret_cast<uint8_t*>(b + oF);
    w.blk = reinterpret_cast<int*>(b + oB);
    w.frame_tot = reinterpret_cast<int*>(b + oT);
    w.total = reinterpret_cast<int*>(b + oS);
    w.bytes = off;
    w.nb = nb;
    return w;
}

static int check_ccl_args(const char* who, const void* a, const void* b, int T, int H, int W, void* ws, size_t ws_bytes,
                          size_t need) {
    if (!a || !b || !ws || T < 0 || H <= 0 || W <= 0) { set_error("%s: invalid argument", who); return TF_ERR_INVALID_ARGUMENT; }
    if ((long long)H * W > 0x7fffffffLL) { set_error("%s: frame too large", who); return TF_ERR_INVALID_ARGUMENT; }
    if (T > 65535) { set_error("%s: more than 65535 frames per call", who); return TF_ERR_UNSUPPORTED; }
    if (ws_bytes < need) { set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, need); return TF_ERR_WORKSPACE_TOO_SMALL; }
    return TF_OK;
}

static void run_ccl(const uint8_t* mask, const CclWs& w, int T, int H, int W, int invert, int conn8, cudaStream_t s) {
    const long long hw = (long long)H * W;
    dim3 blk(32, 8), grid(cdiv(W, 32), cdiv(H, 8), T);
    ccl_init_kernel<<<grid, blk, 0, s>>>(mask, w.L, H, W, invert);
    ccl_merge_kernel<<<grid, blk, 0, s>>>(w.L, H, W, conn8);
    dim3 g1((unsigned)((hw + 255) / 256), T);
    ccl_flatten_kernel<<<g1, 256, 0, s>>>(w.L, hw);
}

}  // namespace tf

using namespace tf;

extern "C" size_t tf_ccl_workspace_bytes(int T, int H, int W) {
    if (T <= 0 || H <= 0 || W <= 0) return 0;
    return carve_ccl(nullptr, T, H, W).bytes;
}

extern "C" int tf_flat_label(const uint8_t* mask, int32_t* labels, int T, int H, int W, int connectivity,
                             int32_t* n_labels, void* workspace, size_t workspace_bytes, void* stream) {
    if (T == 0) return TF_OK;
    const CclWs w = carve_ccl(workspace, T > 0 ? T : 1, H > 0 ? H : 1, W > 0 ? W : 1);
    int rc = check_ccl_args("tf_flat_label", mask, labels, T, H, W, workspace, workspace_bytes, w.bytes);
    if (rc != TF_OK) return rc;
    if (connectivity != 1 && connectivity != 2) { set_error("tf_flat_label: connectivity must be 1 (cross) or 2 (3x3)"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    const long long hw = (long long)H * W;
    LaunchTimer lt(KC_LABEL, (1.0 + 4.0 * 4 + 4.0 * 3) * hw * T, s, 8);
    run_ccl(mask, w, T, H, W, 0, connectivity == 2, s);
    dim3 gn(w.nb, T);
    ccl_count_roots_kernel<<<gn, NUM_BLK, 0, s>>>(w.L, hw, w.nb, w.blk);
    row_exclusive_scan_kernel<<<T, 1024, 0, s>>>(w.blk, w.nb, w.frame_tot);
    row_exclusive_scan_kernel<<<1, 1024, 0, s>>>(w.frame_tot, T, w.total);
    ccl_assign_kernel<<<gn, NUM_BLK, 0, s>>>(w.L, hw, w.nb, w.blk, w.frame_tot, labels);
    dim3 g1((unsigned)((hw + 255) / 256), T);
    ccl_final_kernel<<<g1, 256, 0, s>>>(w.L, hw, labels);
    if (n_labels) store_total_kernel<<<1, 1, 0, s>>>(w.total, n_labels);
    return check_launch("tf_flat_label");
}

extern "C" int tf_binary_fill_holes(const uint8_t* mask, uint8_t* out, int T, int H, int W, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    if (T == 0) return TF_OK;
    const CclWs w = carve_ccl(workspace, T > 0 ? T : 1, H > 0 ? H : 1, W > 0 ? W : 1);
    int rc = check_ccl_args("tf_binary_fill_holes", mask, out, T, H, W, workspace, workspace_bytes, w.bytes);
    if (rc != TF_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const long long hw = (long long)H * W;
    LaunchTimer lt(KC_MORPH, (1.0 + 4.0 * 4 + 4.0 + 2.0 + 1.0) * hw * T, s, 6);
    run_ccl(mask, w, T, H, W, 1, 0, s);
    cudaMemsetAsync(w.flag, 0, (size_t)T * hw, s);
    dim3 gb(cdiv(2 * W + 2 * H, 256), T);
    ccl_border_flag_kernel<<<gb, 256, 0, s>>>(w.L, H, W, w.flag);
    dim3 g1((unsigned)((hw + 255) / 256), T);
    fill_holes_kernel<<<g1, 256, 0, s>>>(w.L, w.flag, hw, out);
    return check_launch("tf_binary_fill_holes");
}
