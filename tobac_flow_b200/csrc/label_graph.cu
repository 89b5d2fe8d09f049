// K13: the label-overlap graph of flow_label / flow_link_overlap (tobac_flow/label.py:84-175, 249-321).
//
// The reference walks every label's pixel list and histograms the labels its pixels land on in the flow-warped
// neighbouring frames (find_neighbour_labels -> utils/label_utils.py:352-376).  Here one pass over the pixels builds
// the same histogram for all labels at once in an open-addressing hash table keyed by (direction, label, neighbour);
// lanes of a warp that hold the same key are merged with match_any before the atomic.  The linking itself (a
// breadth-first walk over a few thousand labels whose result depends on the visiting order) runs on the host in
// tf_label_link_groups, in exactly the reference's order; tf_relabel then writes the final labels.
#include <algorithm>
#include <vector>

#include "tf_common.cuh"

namespace tf {

constexpr unsigned long long kEmptyKey = ~0ull;

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

__device__ __forceinline__ void table_add(unsigned long long* keys, int* counts, long long cap_mask, unsigned long long key,
                                          int c, int* overflow) {
    long long h = (long long)(mix64(key) & (unsigned long long)cap_mask);
    for (long long probes = 0; probes <= cap_mask; ++probes) {
        const unsigned long long prev = atomicCAS(&keys[h], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) { atomicAdd(&counts[h], c); return; }
        h = (h + 1) & cap_mask;
    }
    *overflow = 1;
}

// add `key` (or nothing when !valid) once per distinct key of the warp, weighted by its multiplicity
__device__ __forceinline__ void warp_table_add(unsigned long long* keys, int* counts, long long cap_mask,
                                               unsigned long long key, bool valid, int* overflow) {
    const unsigned grp = __match_any_sync(0xffffffffu, valid ? key : kEmptyKey);
    if (valid && (__ffs(grp) - 1) == (int)(threadIdx.x & 31)) table_add(keys, counts, cap_mask, key, __popc(grp), overflow);
}

__global__ void __launch_bounds__(256) label_overlap_kernel(const int* __restrict__ flat, const int* __restrict__ back,
                                                            const int* __restrict__ fwd, long long n,
                                                            unsigned long long* __restrict__ keys, int* __restrict__ counts,
                                                            long long cap_mask, int* __restrict__ overflow) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (n + 31) / 32 * 32;   // whole warps stay in the loop together (match_any needs them)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        int L = 0, mf = 0, mb = 0;
        if (i < n) {
            L = flat[i];
            if (L > 0) { mf = fwd[i]; mb = back[i]; }
        }
        const unsigned long long kf = ((unsigned long long)(unsigned)L << 31) | (unsigned long long)(unsigned)mf;
        const unsigned long long kb = (1ull << 62) | ((unsigned long long)(unsigned)L << 31) | (unsigned long long)(unsigned)mb;
        warp_table_add(keys, counts, cap_mask, kf, L > 0 && mf > 0, overflow);
        warp_table_add(keys, counts, cap_mask, kb, L > 0 && mb > 0, overflow);
    }
}

// np.bincount(flat.ravel()) for labels 0..n_labels (larger / negative values are an error flagged in `bad`)
__global__ void __launch_bounds__(256) label_sizes_kernel(const int* __restrict__ flat, long long n, int* __restrict__ sizes,
                                                          int n_labels, int* __restrict__ bad) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const int L = i < n ? flat[i] : -1;
        const bool valid = i < n && L >= 0 && L <= n_labels;
        if (i < n && !valid) *bad = 1;
        const unsigned grp = __match_any_sync(0xffffffffu, valid ? L : -1);
        if (valid && L > 0 && (__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sizes[L], __popc(grp));
    }
}

__global__ void __launch_bounds__(256) relabel_kernel(const int* __restrict__ flat, const int* __restrict__ map,
                                                      int* __restrict__ out, long long n, int n_labels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int L = flat[i];
    out[i] = (L > 0 && L <= n_labels) ? map[L] : 0;
}

// per label: first / last frame it appears in and whether it touches mask_a / mask_b anywhere
// (filter_labels_by_length: ndi.find_objects time extent; filter_labels_by_mask: labeled_comprehension(np.any))
__global__ void __launch_bounds__(256) label_stats_kernel(const int* __restrict__ labels, const uint8_t* __restrict__ mask_a,
                                                          const uint8_t* __restrict__ mask_b, long long hw, int n_labels,
                                                          int* __restrict__ tmin, int* __restrict__ tmax,
                                                          int* __restrict__ any_a, int* __restrict__ any_b) {
    const int t = blockIdx.y;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_round = (hw + 31) / 32 * 32;
    const int lane = threadIdx.x & 31;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const long long o = (long long)t * hw + i;
        const int L = i < hw ? labels[o] : 0;
        const bool valid = L > 0 && L <= n_labels;
        const bool a = valid && mask_a && mask_a[o], b = valid && mask_b && mask_b[o];
        const unsigned grp = __match_any_sync(0xffffffffu, valid ? L : 0);
        const unsigned ga = __ballot_sync(0xffffffffu, a) & grp, gb = __ballot_sync(0xffffffffu, b) & grp;
        if (valid && (__ffs(grp) - 1) == lane) {
            atomicMin(&tmin[L], t);
            atomicMax(&tmax[L], t);
            if (ga) any_a[L] = 1;
            if (gb) any_b[L] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) max_label_kernel(const int* __restrict__ flat, long long n, int* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) m = max(m, flat[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

}  // namespace tf

using namespace tf;

extern "C" int tf_label_max(const int32_t* flat, long long n, int32_t* out_max, void* stream) {
    if (!flat || !out_max || n < 0) { set_error("tf_label_max: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(out_max, 0, sizeof(int32_t), s);
    if (n == 0) return TF_OK;
    LaunchTimer lt(KC_LABEL, 4.0 * n, s, 1);
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    max_label_kernel<<<blocks, 256, 0, s>>>(flat, n, out_max);
    return check_launch("tf_label_max");
}

extern "C" int tf_label_stats(const int32_t* labels, const uint8_t* mask_a, const uint8_t* mask_b, int T, long long hw,
                              int n_labels, int32_t* tmin, int32_t* tmax, int32_t* any_a, int32_t* any_b, void* stream) {
    if (!labels || !tmin || !tmax || !any_a || !any_b || T < 0 || hw < 0 || n_labels < 0) {
        set_error("tf_label_stats: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    if (T > 65535) { set_error("tf_label_stats: more than 65535 frames per call"); return TF_ERR_UNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nb = (size_t)(n_labels + 1) * sizeof(int32_t);
    cudaMemsetAsync(tmin, 0x7f, nb, s);    // 0x7f7f7f7f: larger than any frame index
    cudaMemsetAsync(tmax, 0xff, nb, s);    // -1
    cudaMemsetAsync(any_a, 0, nb, s);
    cudaMemsetAsync(any_b, 0, nb, s);
    if (T == 0 || hw == 0) return TF_OK;
    LaunchTimer lt(KC_LABEL, (4.0 + (mask_a ? 1.0 : 0.0) + (mask_b ? 1.0 : 0.0)) * hw * T, s, 1);
    dim3 grid((unsigned)std::min<long long>((hw + 255) / 256, 148 * 8), T);
    label_stats_kernel<<<grid, 256, 0, s>>>(labels, mask_a, mask_b, hw, n_labels, tmin, tmax, any_a, any_b);
    return check_launch("tf_label_stats");
}

extern "C" int tf_label_overlap_count(const int32_t* flat, const int32_t* back, const int32_t* fwd, long long n,
                                      int32_t* sizes, int n_labels, unsigned long long* keys, int32_t* counts,
                                      long long capacity, int32_t* flags, void* stream) {
    if (!flat || !back || !fwd || !sizes || !keys || !counts || !flags || n < 0 || n_labels < 0) {
        set_error("tf_label_overlap_count: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    if (capacity < 64 || (capacity & (capacity - 1))) { set_error("tf_label_overlap_count: capacity must be a power of two >= 64"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(keys, 0xff, (size_t)capacity * sizeof(unsigned long long), s);
    cudaMemsetAsync(counts, 0, (size_t)capacity * sizeof(int32_t), s);
    cudaMemsetAsync(sizes, 0, (size_t)(n_labels + 1) * sizeof(int32_t), s);
    cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), s);
    if (n == 0) return TF_OK;
    LaunchTimer lt(KC_LABEL, 16.0 * n, s, 2);
    const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    label_sizes_kernel<<<blocks, 256, 0, s>>>(flat, n, sizes, n_labels, flags + 1);
    label_overlap_kernel<<<blocks, 256, 0, s>>>(flat, back, fwd, n, keys, counts, capacity - 1, flags);
    return check_launch("tf_label_overlap_count");
}

// Host: the linking walk of flow_label (label.py:139-163) over the histogram entries.  All pointers are HOST pointers.
//   keys/counts   the table copied back from tf_label_overlap_count (empty slots have key ~0)
//   sizes         np.bincount(flat_labels) for 0..n_labels
//   map           out, n_labels + 1 ints: final label of every flat label (map[0] = 0)
// Returns the number of linked objects (>= 0) or a negative tf_status.
extern "C" int tf_label_link_groups(const unsigned long long* keys, const int32_t* counts, long long capacity,
                                    const int32_t* sizes, int n_labels, double overlap, int absolute_overlap,
                                    int32_t* map) {
    if (!keys || !counts || !sizes || !map || n_labels < 0 || capacity < 0) {
        set_error("tf_label_link_groups: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    struct Edge { unsigned long long key; int count; };
    std::vector<Edge> edges;
    for (long long i = 0; i < capacity; ++i)
        if (keys[i] != kEmptyKey) edges.push_back(Edge{keys[i], counts[i]});
    // order: forward entries of a label by ascending neighbour (np.unique), then its backward entries
    std::sort(edges.begin(), edges.end(), [](const Edge& a, const Edge& b) {
        const unsigned long long la = (a.key >> 31) & 0x7fffffffull, lb = (b.key >> 31) & 0x7fffffffull;
        if (la != lb) return la < lb;
        return a.key < b.key;   // direction bit (62) then neighbour
    });
    std::vector<long long> first(n_labels + 2, 0);
    for (const Edge& e : edges) {
        const unsigned long long l = (e.key >> 31) & 0x7fffffffull;
        if (l > (unsigned long long)n_labels) { set_error("tf_label_link_groups: label out of range"); return TF_ERR_INVALID_ARGUMENT; }
        ++first[l + 1];
    }
    for (int l = 0; l <= n_labels; ++l) first[l + 1] += first[l];
    std::vector<char> processed(n_labels + 1, 0);
    std::vector<int> stack;
    int groups = 0;
    map[0] = 0;
    for (int label = 1; label <= n_labels; ++label) {
        if (processed[label]) continue;
        ++groups;
        stack.clear();
        stack.push_back(label);
        processed[label] = 1;
        for (size_t i = 0; i < stack.size(); ++i) {
            const int cur = stack[i];
            const int n_locs = sizes[cur];
            if (n_locs <= 0) continue;   // bins[label] > bins[label - 1]
            for (long long e = first[cur]; e < first[cur + 1]; ++e) {
                const int m = (int)(edges[e].key & 0x7fffffffull);
                const int c = edges[e].count;
                if (m <= 0 || m > n_labels) continue;
                if (!(c > absolute_overlap)) continue;
                if (!((double)c >= overlap * (double)std::min(n_locs, sizes[m]))) continue;
                if (!processed[m]) { stack.push_back(m); processed[m] = 1; }
            }
        }
        for (int l : stack) map[l] = groups;
    }
    return groups;
}

extern "C" int tf_relabel(const int32_t* flat, const int32_t* map, int32_t* out, long long n, int n_labels, void* stream) {
    if (n == 0) return TF_OK;
    if (!flat || !map || !out || n < 0 || n_labels < 0) { set_error("tf_relabel: invalid argument"); return TF_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)stream;
    LaunchTimer lt(KC_LABEL, 8.0 * n, s, 1);
    relabel_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(flat, map, out, n, n_labels);
    return check_launch("tf_relabel");
}
