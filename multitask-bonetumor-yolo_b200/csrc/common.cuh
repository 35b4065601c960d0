// Shared device/host helpers of libbtpost (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/btpost.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbtpost is written for sm_100a (B200) only"
#endif

namespace bt {

// sigmoid(x) > 0.5 in fp32 (src/running_main_v2.py:702-703, src/test_model.py:85) holds exactly
// for x > 1.5 * 2^-24 (pinned against torch in tests/golden/make_golden.py).
__device__ __forceinline__ bool sigmoid_gt_half(float x) { return x > 8.940696716308594e-08f; }

// Workspace carve-up (all offsets 256-byte aligned).
struct Workspace {
    float4 *cand_box;     // [B, cap]
    float *cand_score;    // [B, cap]
    int32_t *cand_label;  // [B, cap]
    int32_t *cand_anchor; // [B, cap]
    unsigned long long *sort_keys;  // [B, cap_pow2] (only used when the list exceeds shared memory)
    int32_t *strip_done;  // [B] strips finished per image (mask kernel)
    int32_t *acc;         // [B, 8] per-image int counters: seg inter,P,G ; uni inter,P,G
    size_t bytes;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline int cand_capacity(const BtParams *p) {
    return (p->max_cand > 0 && p->max_cand < p->num_anchors) ? p->max_cand : p->num_anchors;
}

static inline int next_pow2(int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}

static inline Workspace carve(const BtParams *p, void *base) {
    Workspace w;
    char *ptr = static_cast<char *>(base);
    size_t off = 0;
    const size_t B = (size_t)p->batch, cap = (size_t)cand_capacity(p);
    auto take = [&](size_t n) {
        char *r = ptr ? ptr + off : nullptr;
        off += align_up(n, 256);
        return r;
    };
    w.cand_box = reinterpret_cast<float4 *>(take(B * cap * sizeof(float4)));
    w.cand_score = reinterpret_cast<float *>(take(B * cap * sizeof(float)));
    w.cand_label = reinterpret_cast<int32_t *>(take(B * cap * sizeof(int32_t)));
    w.cand_anchor = reinterpret_cast<int32_t *>(take(B * cap * sizeof(int32_t)));
    w.sort_keys = reinterpret_cast<unsigned long long *>(take(B * (size_t)next_pow2((int)cap) * 8));
    w.strip_done = reinterpret_cast<int32_t *>(take(B * sizeof(int32_t)));
    w.acc = reinterpret_cast<int32_t *>(take(B * 8 * sizeof(int32_t)));
    w.bytes = off;
    return w;
}

int check_params(const BtParams *p, const BtIO *io);
int launch_decode_filter(const BtParams &p, const BtIO &io, const Workspace &w, cudaStream_t s);
int launch_nms_match(const BtParams &p, const BtIO &io, const Workspace &w, cudaStream_t s);
int launch_masks(const BtParams &p, const BtIO &io, const Workspace &w, cudaStream_t s);

}  // namespace bt
