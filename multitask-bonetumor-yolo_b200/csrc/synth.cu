// Device-side generator of the synthetic workload (include/btpost_synth.h): bench / sweep / test infrastructure.
// Bit-identical to btpost/synth.py: splitmix64 finaliser per element, every floating-point step one correctly
// rounded fp32 operation (no fused multiply-adds: this file is compiled with -fmad=false like the rest).
#include "common.cuh"
#include "../../include/btpost_synth.h"

namespace bt {

typedef unsigned long long u64;
constexpr u64 GOLD = 0x9E3779B97F4A7C15ull;
constexpr int HEAD_STREAMS_FIXED = 16, MAX_OBJ = 3;

__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u64 hash_elem(u64 key, u64 idx) { return mix64(key + idx * GOLD); }
__device__ __forceinline__ float u24(u64 h) { return __fmul_rn((float)(unsigned)(h >> 40), 5.9604644775390625e-08f); }
__device__ __forceinline__ float gauss16(u64 h) {
    const int s = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)(h >> 48);
    return __fmul_rn((float)(s - 131070), 2.6428997e-05f);   // np.float32(np.sqrt(3.0) / 65536.0)
}
__device__ __forceinline__ float pow96(float u) {
    const float u2 = __fmul_rn(u, u), u4 = __fmul_rn(u2, u2), u8 = __fmul_rn(u4, u4), u16 = __fmul_rn(u8, u8),
                u32 = __fmul_rn(u16, u16), u64_ = __fmul_rn(u32, u32);
    return __fmul_rn(u64_, u32);
}

// protos[b] = gauss16(hash(key_b, idx)), idx over nm * P * P
__global__ void synth_protos_kernel(const u64 *keys, float *protos, long long per_image) {
    const int b = blockIdx.y;
    const u64 key = keys[b];
    float *o = protos + (size_t)b * per_image;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (long long)gridDim.x * blockDim.x)
        o[i] = gauss16(hash_elem(key, (u64)i));
}

// head rows: one thread per anchor (synth._anchor_fields + make_head_l2)
__global__ void synth_head_kernel(const u64 *keys, const float *objects, float *head, int S, int nc, int nm, int N) {
    const int b = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const u64 key = keys[b];
    const int KS = HEAD_STREAMS_FIXED + nc + nm;
    const u64 base = (u64)n * (u64)KS;
    auto U = [&](int s) { return u24(hash_elem(key, base + (u64)s)); };
    // anchor centre: levels P3, P4, P5, row-major
    float ax, ay;
    {
        int m = n, st = 8;
        for (; st <= 32; st <<= 1) {
            const int w = S / st, cnt = w * (S / st);
            if (m < cnt || st == 32) {
                const int y = m / w, x = m - y * w;
                ax = __fmul_rn(__fadd_rn((float)x, 0.5f), (float)st);
                ay = __fmul_rn(__fadd_rn((float)y, 0.5f), (float)st);
                break;
            }
            m -= cnt;
        }
    }
    float cx = ax, cy = ay;
    float w = __fadd_rn(4.0f, __fmul_rn(60.0f, U(1))), h = __fadd_rn(4.0f, __fmul_rn(60.0f, U(2)));
    const bool coin = U(0) < 0.5f;
    const float j0 = __fadd_rn(__fadd_rn(U(3), U(4)), -1.0f), j1 = __fadd_rn(__fadd_rn(U(5), U(6)), -1.0f);
    const float j2 = __fadd_rn(__fadd_rn(U(7), U(8)), -1.0f), j3 = __fadd_rn(__fadd_rn(U(9), U(10)), -1.0f);
    const float s_obj = __fadd_rn(0.3f, __fmul_rn(0.65f, U(11)));
    int obj_cls = -1;
    bool assigned = false;
    for (int o = 0; o < MAX_OBJ; ++o) {
        const float *t = objects + ((size_t)b * MAX_OBJ + o) * 6;
        if (t[0] == 0.0f) continue;
        const float ocx = t[2], ocy = t[3], ow = t[4], oh = t[5];
        const bool inside = fabsf(__fadd_rn(ax, -ocx)) < __fmul_rn(ow, 0.5f) && fabsf(__fadd_rn(ay, -ocy)) < __fmul_rn(oh, 0.5f) &&
                            coin && !assigned;
        if (inside) {
            const float sw = __fmul_rn(0.06f, ow), sh = __fmul_rn(0.06f, oh);
            cx = __fadd_rn(ocx, __fmul_rn(j0, sw));
            cy = __fadd_rn(ocy, __fmul_rn(j1, sh));
            w = __fadd_rn(ow, __fmul_rn(j2, sw));
            h = __fadd_rn(oh, __fmul_rn(j3, sh));
            obj_cls = (int)t[1];
            assigned = true;
        }
    }
    float *hb = head + (size_t)b * (4 + nc + nm) * N + n;
    hb[0] = cx; hb[(size_t)N] = cy; hb[(size_t)2 * N] = w; hb[(size_t)3 * N] = h;
    for (int c = 0; c < nc; ++c) {
        const float u = U(HEAD_STREAMS_FIXED + c);
        float sc = __fadd_rn(0.002f, __fmul_rn(0.6f, pow96(u)));
        if (assigned) sc = (c == obj_cls) ? s_obj : __fmul_rn(0.02f, u);
        hb[(size_t)(4 + c) * N] = sc;
    }
    for (int m = 0; m < nm; ++m) hb[(size_t)(4 + nc + m) * N] = gauss16(hash_elem(key, base + (u64)(HEAD_STREAMS_FIXED + nc + m)));
}

// GT mask: union of the ellipses inscribed in the object boxes
__global__ void synth_masks_kernel(const float *objects, uint8_t *masks, int S) {
    const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= S) return;
    const float px = __fadd_rn((float)x, 0.5f), py = __fadd_rn((float)y, 0.5f);
    unsigned v = 0;
    for (int o = 0; o < MAX_OBJ; ++o) {
        const float *t = objects + ((size_t)b * MAX_OBJ + o) * 6;
        if (t[0] == 0.0f) continue;
        const float dx = __fdiv_rn(__fadd_rn(px, -t[2]), __fmul_rn(t[4], 0.5f));
        const float dy = __fdiv_rn(__fadd_rn(py, -t[3]), __fmul_rn(t[5], 0.5f));
        if (__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx)) <= 1.0f) v = 1;
    }
    masks[((size_t)b * S + y) * S + x] = (uint8_t)v;
}

}  // namespace bt

using namespace bt;

extern "C" int btpost_synth_batch(int32_t batch, int32_t img_size, int32_t nc, int32_t nm, const uint64_t *keys_head,
                                  const uint64_t *keys_proto, const float *objects, float *head, float *protos,
                                  uint8_t *masks_gt, void *stream) {
    if (batch <= 0 || img_size <= 0 || img_size % 32 != 0 || nc <= 0 || nm <= 0) return BT_ERR_BAD_ARG;
    if (batch > 65535 || img_size > 65535) return BT_ERR_UNSUPPORTED;
    if ((head || masks_gt) && !objects) return BT_ERR_BAD_ARG;
    if ((head && !keys_head) || (protos && !keys_proto)) return BT_ERR_BAD_ARG;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int S = img_size, P = S / 4;
    int N = 0;
    for (int st = 8; st <= 32; st *= 2) N += (S / st) * (S / st);
    if (protos) {
        const long long per = (long long)nm * P * P;
        synth_protos_kernel<<<dim3(296, batch), 256, 0, s>>>(reinterpret_cast<const u64 *>(keys_proto), protos, per);
    }
    if (head)
        synth_head_kernel<<<dim3((N + 127) / 128, batch), 128, 0, s>>>(reinterpret_cast<const u64 *>(keys_head), objects, head, S, nc, nm, N);
    if (masks_gt) synth_masks_kernel<<<dim3((S + 255) / 256, S, batch), 256, 0, s>>>(objects, masks_gt, S);
    return cudaGetLastError() == cudaSuccess ? BT_OK : BT_ERR_CUDA;
}
