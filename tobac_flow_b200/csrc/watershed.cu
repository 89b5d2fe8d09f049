// Semi-Lagrangian watershed flood (tobac_flow/_watershed.pyx:222-344, called from tobac_flow/watershed.py:17-168).
//
// The algorithm is a priority flood from the marker pixels: the queue is ordered by (field value, push age), a pixel takes
// the label of the pixel that pushed it at push time, and the neighbours of a pixel in the adjacent time steps are
// displaced by the rounded optical-flow vector at that pixel (forward flow for the t + 1 neighbours, backward flow for
// the t - 1 neighbours).  The label of a pixel depends on the global pop order (ties between basins are decided by push
// age), so the flood is inherently sequential: like the reference's Cython, it runs on the host, inside the library, on
// host buffers (the flow-offset fields it consumes are prepared on the device by tobac_flow_b200/watershed.py).
// Only the plain watershed of the reference's call site is built (compactness = 0, no watershed lines).
#include <vector>

#include "tf_common.cuh"

namespace tf {

struct WsItem {
    float value;
    int age;
    long long index;
};

// the reference's ordering (_watershed.pyx:161-164): by value, ties by age (ages are unique)
static inline bool ws_smaller(const WsItem& a, const WsItem& b) {
    if (a.value != b.value) return a.value < b.value;
    return a.age < b.age;
}

// binary min-heap with the reference's sift rules (append + sift up; move last to root + sift down towards the smaller
// child), so that even fields with NaNs -- for which (value, age) is not a strict weak order -- pop in the same sequence
struct WsHeap {
    std::vector<WsItem> h;
    void push(const WsItem& e) {
        h.push_back(e);
        size_t child = h.size() - 1;
        while (child > 0) {
            const size_t parent = (child + 1) / 2 - 1;
            if (!ws_smaller(h[child], h[parent])) break;
            std::swap(h[child], h[parent]);
            child = parent;
        }
    }
    WsItem pop() {
        const WsItem top = h[0];
        const size_t n = h.size() - 1;
        if (n == 0) { h.pop_back(); return top; }
        h[0] = h[n];
        h.pop_back();
        size_t i = 0;
        for (;;) {
            const size_t l = 2 * i + 1, r = 2 * i + 2;
            if (l >= n) break;
            size_t smallest = i;
            if (ws_smaller(h[l], h[i])) smallest = l;
            if (r < n && ws_smaller(h[r], h[smallest])) smallest = r;
            if (smallest == i) break;
            std::swap(h[i], h[smallest]);
            i = smallest;
        }
        return top;
    }
};

}  // namespace tf

using namespace tf;

extern "C" int tf_watershed_flood_host(const float* image, const long long* marker_locations, long long n_markers,
                                       const long long* structure, int n_neighbors, const int* forward_offset,
                                       const int* backward_offset, const int* forward_offset_locations,
                                       const int* backward_offset_locations, const signed char* mask, int* output,
                                       long long n) {
    if (!image || !structure || !forward_offset || !backward_offset || !forward_offset_locations ||
        !backward_offset_locations || !mask || !output || n <= 0 || n_neighbors <= 0 || n_markers < 0 ||
        (n_markers > 0 && !marker_locations)) {
        set_error("tf_watershed_flood_host: invalid argument");
        return TF_ERR_INVALID_ARGUMENT;
    }
    WsHeap heap;
    heap.h.reserve(1024);
    for (long long i = 0; i < n_markers; ++i) {
        const long long index = marker_locations[i];
        if (index < 0 || index >= n) { set_error("tf_watershed_flood_host: marker location out of range"); return TF_ERR_INVALID_ARGUMENT; }
        heap.push(WsItem{image[index], 0, index});
    }
    long long age = 1;
    while (!heap.h.empty()) {
        const WsItem e = heap.pop();
        for (int i = 0; i < n_neighbors; ++i) {
            const long long nb = structure[i] + e.index + (long long)forward_offset_locations[i] * forward_offset[e.index] +
                                 (long long)backward_offset_locations[i] * backward_offset[e.index];
            // the padded border (mask == 0 there) keeps every neighbour inside the volume; the range test only guards
            // against inconsistent inputs
            if (nb < 0 || nb >= n) continue;
            if (!mask[nb]) continue;          // outside the mask (includes the padding)
            if (output[nb]) continue;         // already labelled
            ++age;
            output[nb] = output[e.index];     // plain watershed: label at push time
            heap.push(WsItem{image[nb], (int)age, nb});
        }
    }
    return TF_OK;
}
