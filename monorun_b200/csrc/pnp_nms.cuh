// pnp_nms.cuh -- class-wise 3-D NMS on bird's-eye-view rotated boxes: MonoRUnRoIHead.multiclass_3d_result_nms
// (monorun_roi_head.py:619-655): per class, boxes in descending score order, a box is dropped if its rotated BEV IoU
// with an earlier KEPT box of the same class exceeds nms_thr (mmdet3d.ops.iou3d.nms_gpu on
// xywhr2xyxyr(bbox_3d[:, [3, 5, 0, 2, 6]]), i.e. centre (x, z), extent (l, w), angle ry; :660-680).
//
// One CTA per image (group of objects): bitonic sort of (score, index) in shared memory, then the greedy scan in
// sorted order -- for every surviving box all threads test the later boxes in parallel (no N x N mask in memory).
// The intersection area of two rotated rectangles comes from clipping one against the four half-planes of the other
// (Sutherland-Hodgman), fp32.
#pragma once
#include "pnp_device.cuh"

namespace mrpnp {

constexpr int kNmsThreads = 256;

struct BevBox {
    float cx, cz, hl, hw, c, s;  // centre, half extents, cos / sin of ry
};

__device__ __forceinline__ float bev_intersection(const BevBox& a, const BevBox& b) {
    // corners of a, expressed in b's frame (b becomes the axis-aligned rectangle [-hl, hl] x [-hw, hw])
    float px[8], pz[8], qx[8], qz[8];
    const float dx = a.cx - b.cx, dz = a.cz - b.cz;
    // relative rotation: R_b^T R_a ; world offsets of a's corners: (c_a u + s_a v, -s_a u + c_a v)
    const float ux[4] = {a.hl, -a.hl, -a.hl, a.hl}, vz[4] = {a.hw, a.hw, -a.hw, -a.hw};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float wx = dx + a.c * ux[k] + a.s * vz[k], wz = dz - a.s * ux[k] + a.c * vz[k];
        px[k] = b.c * wx - b.s * wz;   // inverse of (c u + s v, -s u + c v)
        pz[k] = b.s * wx + b.c * wz;
    }
    int n = 4;
    // clip against x <= hl, x >= -hl, z <= hw, z >= -hw
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float lim = (e < 2) ? b.hl : b.hw;
        const float sgn = (e & 1) ? -1.f : 1.f;
        int m = 0;
        for (int k = 0; k < n; ++k) {
            const int k2 = (k + 1 == n) ? 0 : k + 1;
            const float ck = ((e < 2) ? px[k] : pz[k]) * sgn, ck2 = ((e < 2) ? px[k2] : pz[k2]) * sgn;
            const bool in1 = ck <= lim, in2 = ck2 <= lim;
            if (in1) { qx[m] = px[k]; qz[m] = pz[k]; ++m; }
            if (in1 != in2) {
                const float t = (lim - ck) / (ck2 - ck);
                qx[m] = px[k] + t * (px[k2] - px[k]);
                qz[m] = pz[k] + t * (pz[k2] - pz[k]);
                ++m;
            }
        }
        n = m;
        for (int k = 0; k < n; ++k) { px[k] = qx[k]; pz[k] = qz[k]; }
        if (n == 0) return 0.f;
    }
    float area = 0.f;
    for (int k = 0; k < n; ++k) {
        const int k2 = (k + 1 == n) ? 0 : k + 1;
        area += px[k] * pz[k2] - px[k2] * pz[k];
    }
    return 0.5f * fabsf(area);
}

__device__ __forceinline__ float bev_iou(const BevBox& a, const BevBox& b) {
    const float inter = bev_intersection(a, b);
    const float ua = 4.f * a.hl * a.hw, ub = 4.f * b.hl * b.hw;
    return inter / fmaxf(ua + ub - inter, 1e-8f);
}

// shared memory per CTA: cap x (key score, key index, class, removed) + cap x BevBox
__global__ void __launch_bounds__(kNmsThreads) nms_bev_kernel(const float* __restrict__ bbox3d, const long long* __restrict__ labels,
                                                              const int* __restrict__ offsets, float thr, int cap,
                                                              unsigned char* __restrict__ keep) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float* key = reinterpret_cast<float*>(nms_smem);
    int* idx = reinterpret_cast<int*>(key + cap);
    int* cls = idx + cap;
    int* removed = cls + cap;
    BevBox* box = reinterpret_cast<BevBox*>(removed + cap);
    const int g = blockIdx.x, tid = threadIdx.x;
    const int lo = offsets[g], n = offsets[g + 1] - lo;
    if (n <= 0) return;
    int m = 1;
    while (m < n) m <<= 1;  // bitonic size (<= cap, checked by the host)
    for (int i = tid; i < m; i += kNmsThreads) {
        key[i] = i < n ? bbox3d[(size_t)(lo + i) * 8 + 7] : -3.0e38f;
        idx[i] = i < n ? i : 0x7fffffff;
    }
    __syncthreads();
    // descending by score, ascending by index among equal scores (the order of a stable descending sort)
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < m; i += kNmsThreads) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const float ki = key[i], kl = key[l];
                    const int ii = idx[i], il = idx[l];
                    const bool i_first = (ki > kl) || (ki == kl && ii < il);   // i belongs before l
                    if (up != i_first) { key[i] = kl; key[l] = ki; idx[i] = il; idx[l] = ii; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += kNmsThreads) {
        const float* b = bbox3d + (size_t)(lo + idx[i]) * 8;   // l,h,w,x,y,z,ry,score
        BevBox q;
        q.cx = b[3]; q.cz = b[5]; q.hl = 0.5f * b[0]; q.hw = 0.5f * b[2];
        sincosf(b[6], &q.s, &q.c);
        box[i] = q;
        cls[i] = labels ? (int)labels[lo + idx[i]] : 0;
        removed[i] = 0;
    }
    __syncthreads();
    for (int a = 0; a < n; ++a) {
        if (!removed[a]) {   // uniform: written before the barrier below
            const BevBox qa = box[a];
            const int ca = cls[a];
            for (int b = a + 1 + tid; b < n; b += kNmsThreads)
                if (!removed[b] && cls[b] == ca && bev_iou(qa, box[b]) > thr) removed[b] = 1;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += kNmsThreads) keep[lo + idx[i]] = removed[i] ? 0 : 1;
}

}  // namespace mrpnp
