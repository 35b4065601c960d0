// Kernel 2: per-image stable descending sort + greedy NMS with early exit at max_det + gather of
// the kept detections + COCOeval per-image matching.
//
// Reference statements replaced (paths under /root/reference/src):
//   NMS       running_main_v2.py:817  torchvision.ops.nms(item_boxes, item_top_scores, NMS_IOU)[:TOP_K]
//             (running_main_v3.py:549).  Keep indices are bit-exact: same ordering (stable
//             descending, NaN first, ties -> lower index), same fp32 IoU expression
//             inter / (area_i + area_j - inter) with one rounding per operation, same strict
//             comparison against the double threshold.
//   package   running_main_v2.py:818-839  ([K,4] boxes, [K] scores, [K] labels, [K,6] log rows)
//   matching  torchmetrics MeanAveragePrecision.update -> pycocotools COCOeval.evaluateImg
//             (configured at running_main_v2.py:228-251, evaluate_model.py:81-94)
//
// B200 mapping: one 1024-thread CTA per image.  The sorted candidate list is consumed in chunks
// of 64: (A) the chunk is tested against the boxes kept so far (kept list staged in shared
// memory, 16 threads per candidate, warp ballot to merge), (B) a 64x64 intra-chunk suppression
// bitmask is built with warp ballots, (C) one warp sweeps the bitmask as a fixpoint (rows in
// registers, warp-wide OR reductions) instead of one dependent load per keep.  Because keeps are
// emitted in score order, `[:TOP_K]` is an early exit: at most max_det * M IoUs are evaluated
// instead of M^2/2, which is what makes the 30k-candidate configuration tractable.
#include "common.cuh"

namespace bt {

constexpr int K2_THREADS = 1024;
constexpr int K2_WARPS = K2_THREADS / 32;
constexpr int NMS_CHUNK = 64;
constexpr int SORT_SMEM_MAX = 8192;  // 64-bit keys sorted in shared memory up to this many

struct K2Params {
    int N, nc, nm, C, cap, cap_pow2, max_det, max_gt, class_mode, layout;
    float thr_up;     // smallest float f with (double)f > iou_thres:  (double)ovr > thr  <=>  ovr >= thr_up
    int early_out;    // iou_thres >= 0: pairs with zero intersection can never be suppressed
    float max_wh;
    const float *head;    // L2: coefficients are rows 4+nc.. of the head
    const float *coeffs;  // L1: [B, nm, N]
    const float4 *cand_box;
    const float *cand_score;
    const int32_t *cand_label, *cand_anchor, *n_cand;
    unsigned long long *sort_keys;
    int32_t *det_count;
    float *dets;
    int64_t *det_keep;
    int32_t *det_anchor;
    float *det_coeff;
    // GT + COCO
    const int32_t *gt_count;
    const float *gt_boxes;
    const int32_t *gt_labels;
    int T;
    double thrs[BT_MAX_IOU_THRS];
    int32_t *dt_match;
    uint8_t *dt_ignore, *gt_ignore;
    // accumulators of the mask kernel, zeroed here
    int32_t *strip_done, *acc, *inst_area, *inst_inter;
    int smem_keys;  // number of 64-bit slots in the shared key region
    int win;        // sorted-candidate window staged in shared memory (multiple of NMS_CHUNK)
    int crop, PW, PH; float rx, ry; short4 *det_region;   // crop regions for the mask kernel
    int centre_cull; // iou_thres >= 0.55: a pair can only suppress if the later box's centre lies in the earlier box
};

__device__ __forceinline__ uint32_t desc_key(float s) {
    uint32_t u = __float_as_uint(s);
    if (s != s) return 0u;                                   // NaN sorts first
    if (u == 0x80000000u) u = 0u;                            // -0.0 == +0.0
    uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;                                             // smaller key = higher score
}

// i = earlier (kept) box, j = later candidate; operand order of std::max/std::min as in
// torchvision's nms_kernel_impl so NaN coordinates behave identically.
//
// centre_cull: IoU > 0.5 means the intersection covers more than half of EACH box, hence contains
// each box's centre.  With iou_thres >= 0.55 (margin >> fp32 rounding of the IoU) a pair whose
// later centre (cxj, cyj) is outside the earlier box can be skipped without evaluating the IoU;
// every pair that is evaluated uses the exact expression, so the keep set does not change.
__device__ __forceinline__ bool suppresses(const float4 &bi, float ai, int li, const float4 &bj, float aj, int lj,
                                           float thr_up, int early_out, int class_mode, int centre_cull = 0,
                                           float cxj = 0.0f, float cyj = 0.0f) {
    if (class_mode == BT_CLASS_AWARE && li != lj) return false;
    if (centre_cull && !(cxj >= bi.x && cxj <= bi.z && cyj >= bi.y && cyj <= bi.w)) return false;
    float xx1 = (bi.x < bj.x) ? bj.x : bi.x;
    float yy1 = (bi.y < bj.y) ? bj.y : bi.y;
    float xx2 = (bj.z < bi.z) ? bj.z : bi.z;
    float yy2 = (bj.w < bi.w) ? bj.w : bi.w;
    float w = __fsub_rn(xx2, xx1), h = __fsub_rn(yy2, yy1);
    w = (0.0f < w) ? w : 0.0f;
    h = (0.0f < h) ? h : 0.0f;
    float inter = __fmul_rn(w, h);
    if (early_out && !(inter > 0.0f)) return false;
    float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
    return ovr >= thr_up;
}

__device__ __forceinline__ double bb_iou(const double *d, const double *g) {
    double w = fmin(d[0] + d[2], g[0] + g[2]) - fmax(d[0], g[0]);
    if (w <= 0) return 0.0;
    double h = fmin(d[1] + d[3], g[1] + g[3]) - fmax(d[1], g[1]);
    if (h <= 0) return 0.0;
    double i = w * h;
    return i / (d[2] * d[3] + g[2] * g[3] - i);
}

__global__ void __launch_bounds__(K2_THREADS) nms_match_kernel(const __grid_constant__ K2Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int K = P.max_det;

    // ---- shared-memory carve-up
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem_raw);          // [smem_keys]
    float4 *s_kbox = reinterpret_cast<float4 *>(s_keys + P.smem_keys);                       // [K]   kept boxes
    float4 *s_sbox = s_kbox + K;                                                             // [win] sorted boxes (window)
    float *s_karea = reinterpret_cast<float *>(s_sbox + P.win);                              // [K]
    float *s_sarea = s_karea + K;                                                            // [win]
    int *s_klabel = reinterpret_cast<int *>(s_sarea + P.win);                                // [K]
    int *s_kidx = s_klabel + K;                                                              // [K]
    int *s_slabel = s_kidx + K;                                                              // [win]
    __shared__ unsigned int s_mask32[NMS_CHUNK * 2];
    __shared__ unsigned int s_supA[2];
    __shared__ unsigned long long s_keepm;
    __shared__ int s_warpcnt[K2_WARPS];

    BT_PHASE_INIT();
    // ---- zero the per-image accumulators the mask kernel adds into
    if (tid < 8) P.acc[b * 8 + tid] = 0;
    if (tid == 8) P.strip_done[b] = 0;
    for (int i = tid; i < K; i += K2_THREADS) {
        if (P.inst_area) P.inst_area[(size_t)b * K + i] = 0;
        if (P.inst_inter) P.inst_inter[(size_t)b * K + i] = 0;
    }

    const int M = P.n_cand[b];
    const float4 *cbox = P.cand_box + (size_t)b * P.cap;
    const float *cscore = P.cand_score + (size_t)b * P.cap;
    const int32_t *clabel = P.cand_label + (size_t)b * P.cap;
    const int32_t *canchor = P.cand_anchor + (size_t)b * P.cap;

    // ---- 1. stable descending sort: key = (descending score key << 32) | candidate index
    int P2 = 1;
    while (P2 < M) P2 <<= 1;
    unsigned long long *keys = (P2 <= P.smem_keys) ? s_keys : (P.sort_keys + (size_t)b * P.cap_pow2);
    for (int i = tid; i < P2; i += K2_THREADS)
        keys[i] = (i < M) ? (((unsigned long long)desc_key(cscore[i]) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P2 >> 1); t += K2_THREADS) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i | j;
                unsigned long long a = keys[i], c = keys[l];
                bool asc = (i & k) == 0;
                if ((a > c) == asc) { keys[i] = c; keys[l] = a; }
            }
            __syncthreads();
        }

    BT_PHASE_MARK(1, 0);   // sort
    // ---- 2. chunked greedy NMS over the sorted list, staged through shared memory in windows
    int nkept = 0;
    for (int c0 = 0; c0 < M && nkept < K; c0 += NMS_CHUNK) {
        const int n_in = min(NMS_CHUNK, M - c0);
        const int w0 = (c0 / P.win) * P.win;      // window holding this chunk (win is a multiple of 64)
        if (c0 == w0) {
            // gather the next window of sorted candidates: boxes (class-offset if asked), areas, labels
            const int wn = min(P.win, M - w0);
            for (int t = tid; t < wn; t += K2_THREADS) {
                const int idx = (int)(unsigned)keys[w0 + t];
                float4 bx = cbox[idx];
                const int lb = clabel[idx];
                if (P.class_mode == BT_CLASS_OFFSET) {
                    const float off = __fmul_rn((float)lb, P.max_wh);
                    bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off);
                    bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
                }
                s_sbox[t] = bx;
                s_sarea[t] = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
                s_slabel[t] = lb;
            }
        }
        if (tid < 2) s_supA[tid] = 0;
        __syncthreads();
        const float4 *cb = s_sbox + (c0 - w0);
        const float *ca = s_sarea + (c0 - w0);
        const int *cl = s_slabel + (c0 - w0);
        // (A) chunk vs boxes kept so far: 16 threads per candidate
        {
            const int ci = tid >> 4, sub = tid & 15;
            bool f = false;
            if (ci < n_in) {
                const float4 bj = cb[ci];
                const float aj = ca[ci];
                const int lj = cl[ci];
                const float cxj = __fmul_rn(__fadd_rn(bj.x, bj.z), 0.5f), cyj = __fmul_rn(__fadd_rn(bj.y, bj.w), 0.5f);
                for (int j = sub; j < nkept; j += 16)
                    f |= suppresses(s_kbox[j], s_karea[j], s_klabel[j], bj, aj, lj, P.thr_up, P.early_out, P.class_mode,
                                    P.centre_cull, cxj, cyj);
            }
            unsigned m = __ballot_sync(0xffffffffu, f);
            if ((lane & 15) == 0 && ((m >> lane) & 0xffffu)) atomicOr(&s_supA[ci >> 5], 1u << (ci & 31));
        }
        BT_PHASE_MARK(1, 8);   // chunk: load + phase A
        // (B) intra-chunk 64x64 upper-triangular suppression bitmask
#pragma unroll
        for (int r = 0; r < (NMS_CHUNK * NMS_CHUNK) / K2_THREADS; ++r) {
            const int q = r * K2_THREADS + tid;
            const int i = q >> 6, j = q & 63;
            bool f = false;
            if (j > i && j < n_in) {
                const float4 bj = cb[j];
                f = suppresses(cb[i], ca[i], cl[i], bj, ca[j], cl[j], P.thr_up, P.early_out, P.class_mode, P.centre_cull,
                               __fmul_rn(__fadd_rn(bj.x, bj.z), 0.5f), __fmul_rn(__fadd_rn(bj.y, bj.w), 0.5f));
            }
            unsigned m = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_mask32[i * 2 + (j >> 5)] = m;
        }
        __syncthreads();
        BT_PHASE_MARK(1, 9);   // chunk: phase B
        // (C) warp-level suppression sweep as a fixpoint: a candidate no undecided earlier candidate
        // can suppress is final; its row removes its victims.  Resolves sparse chunks in 1-3 rounds
        // instead of one dependent shared-memory load per keep.
        if (wid == 0) {
            const unsigned long long row_a = ((unsigned long long)s_mask32[lane * 2 + 1] << 32) | s_mask32[lane * 2];
            const unsigned long long row_b = ((unsigned long long)s_mask32[(lane + 32) * 2 + 1] << 32) | s_mask32[(lane + 32) * 2];
            const unsigned long long valid = (n_in == 64) ? ~0ull : ((1ull << n_in) - 1ull);
            unsigned long long und = valid & ~(((unsigned long long)s_supA[1] << 32) | s_supA[0]);
            unsigned long long keepm = 0ull;
            while (und) {
                const bool ua = (und >> lane) & 1ull, ub = (und >> (lane + 32)) & 1ull;
                const unsigned long long r = (ua ? row_a : 0ull) | (ub ? row_b : 0ull);
                const unsigned long long S = ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(r >> 32)) << 32) |
                                             __reduce_or_sync(0xffffffffu, (unsigned)r);
                const unsigned long long def = und & ~S;
                keepm |= def;
                const bool da = (def >> lane) & 1ull, db = (def >> (lane + 32)) & 1ull;
                const unsigned long long d = (da ? row_a : 0ull) | (db ? row_b : 0ull);
                const unsigned long long Dm = ((unsigned long long)__reduce_or_sync(0xffffffffu, (unsigned)(d >> 32)) << 32) |
                                              __reduce_or_sync(0xffffffffu, (unsigned)d);
                und &= ~(def | Dm);
            }
            // [:TOP_K]: only the first `room` keeps survive (later ones cannot affect earlier ones)
            int room = K - nkept;
            if (__popcll(keepm) > room) {
                unsigned long long t = keepm, kept = 0ull;
                for (int i = 0; i < room; ++i) { unsigned long long low = t & (~t + 1ull); kept |= low; t ^= low; }
                keepm = kept;
            }
            if (lane == 0) s_keepm = keepm;
        }
        __syncthreads();
        BT_PHASE_MARK(1, 10);  // chunk: phase C
        const unsigned long long keepm = s_keepm;
        if (tid < NMS_CHUNK && ((keepm >> tid) & 1ull)) {
            int slot = nkept + __popcll(keepm & ((1ull << tid) - 1ull));
            s_kbox[slot] = cb[tid];
            s_karea[slot] = ca[tid];
            s_klabel[slot] = cl[tid];
            s_kidx[slot] = (int)(unsigned)keys[c0 + tid];
        }
        nkept += __popcll(keepm);
        __syncthreads();
    }

    BT_PHASE_MARK(1, 1);   // NMS chunks
    // ---- 3. package kept detections (running_main_v2.py:818-839); zero-fill the padding
    if (tid == 0) P.det_count[b] = nkept;
    for (int k = tid; k < K; k += K2_THREADS) {
        float *o = P.dets + ((size_t)b * K + k) * 6;
        if (k < nkept) {
            int idx = s_kidx[k];
            float4 bx = cbox[idx];
            o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
            o[4] = cscore[idx];
            o[5] = (float)clabel[idx];
            P.det_keep[(size_t)b * K + k] = idx;
            P.det_anchor[(size_t)b * K + k] = canchor[idx];
            int r_lo, r_hi, c_lo, c_hi;
            const float bb[4] = {bx.x, bx.y, bx.z, bx.w};
            const bool ok = crop_region(bb, P.crop, P.rx, P.ry, P.PW, P.PH, r_lo, r_hi, c_lo, c_hi);
            P.det_region[(size_t)b * K + k] = ok ? make_short4((short)r_lo, (short)r_hi, (short)c_lo, (short)c_hi)
                                                 : make_short4(1, 0, 1, 0);
        } else {
            P.det_region[(size_t)b * K + k] = make_short4(1, 0, 1, 0);
            o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f;
            P.det_keep[(size_t)b * K + k] = -1;
            P.det_anchor[(size_t)b * K + k] = -1;
        }
    }
    for (int q = tid; q < K * 32; q += K2_THREADS) {   // nm == 32 (validated by check_params)
        int k = q >> 5, m = q & 31;
        float v = 0.0f;
        if (k < nkept) {
            int a = canchor[s_kidx[k]];
            v = (P.layout == BT_LAYOUT_L2) ? __ldg(P.head + ((size_t)b * P.C + 4 + P.nc + m) * P.N + a)
                                           : __ldg(P.coeffs + ((size_t)b * P.nm + m) * P.N + a);
        }
        P.det_coeff[((size_t)b * K + k) * 32 + m] = v;
    }

    BT_PHASE_MARK(1, 2);   // package
    // ---- 4. COCOeval.evaluateImg for every (class, area range, IoU threshold)
    if (P.dt_match == nullptr) return;
    __syncthreads();  // key region is free from here on: reuse it for the double-precision tables
    const int G = P.gt_count[b];
    const int D = nkept;
    const int T = P.T;
    double *s_db = reinterpret_cast<double *>(smem_raw);   // [K][4] x, y, w, h
    double *s_gb = s_db + (size_t)K * 4;                   // [max_gt][4]
    double *s_iou = s_gb + (size_t)P.max_gt * 4;           // [D][G] if it fits
    const int iou_room = P.smem_keys - K * 4 - P.max_gt * 4;
    const bool iou_in_smem = (D * G) <= iou_room;
    int *s_dl = s_klabel;                                   // det labels (already there, K entries)
    __shared__ int s_gl[32];
    __shared__ unsigned s_gign[BT_NUM_AREA];
    float *s_maxiou = s_karea;                               // [K] reuse: max IoU over same-class GT (as float of double, rounded up)
    int *s_list = s_kidx;                                    // [K] reuse: compacted list of dets that can match anything
    __shared__ int s_nlist;

    for (int k = tid; k < D; k += K2_THREADS) {
        const float *o = P.dets + ((size_t)b * K + k) * 6;
        s_db[k * 4 + 0] = (double)o[0];
        s_db[k * 4 + 1] = (double)o[1];
        s_db[k * 4 + 2] = (double)__fsub_rn(o[2], o[0]);
        s_db[k * 4 + 3] = (double)__fsub_rn(o[3], o[1]);
        s_dl[k] = (int)o[5];
    }
    if (tid < G) {
        const float *g = P.gt_boxes + ((size_t)b * P.max_gt + tid) * 4;
        s_gb[tid * 4 + 0] = (double)g[0];
        s_gb[tid * 4 + 1] = (double)g[1];
        s_gb[tid * 4 + 2] = (double)__fsub_rn(g[2], g[0]);
        s_gb[tid * 4 + 3] = (double)__fsub_rn(g[3], g[1]);
        s_gl[tid] = P.gt_labels[(size_t)b * P.max_gt + tid];
    }
    __syncthreads();
    const double area_lo[BT_NUM_AREA] = {0.0, 0.0, 32.0 * 32.0, 96.0 * 96.0};
    const double area_hi[BT_NUM_AREA] = {1e10, 32.0 * 32.0, 96.0 * 96.0, 1e10};
    if (tid < BT_NUM_AREA) {
        unsigned m = 0;
        for (int g = 0; g < G; ++g) {
            double ar = s_gb[g * 4 + 2] * s_gb[g * 4 + 3];
            bool in_cls = s_gl[g] >= 0 && s_gl[g] < P.nc;
            bool ig = in_cls && (ar < area_lo[tid] || ar > area_hi[tid]);
            if (ig) m |= 1u << g;
            if (P.gt_ignore) P.gt_ignore[((size_t)b * BT_NUM_AREA + tid) * P.max_gt + g] = ig ? 1 : 0;
        }
        for (int g = G; g < P.max_gt; ++g)
            if (P.gt_ignore) P.gt_ignore[((size_t)b * BT_NUM_AREA + tid) * P.max_gt + g] = 0;
        s_gign[tid] = m;
    }
    // IoU table + per-detection best IoU over same-class GT
    double thr_min = 1.0;
    for (int t = 0; t < T; ++t) thr_min = fmin(thr_min, fmin(P.thrs[t], 1 - 1e-10));
    for (int k = tid; k < D; k += K2_THREADS) {
        double best = -1.0;
        for (int g = 0; g < G; ++g) {
            double v = bb_iou(s_db + k * 4, s_gb + g * 4);
            if (iou_in_smem) s_iou[k * G + g] = v;
            if (s_gl[g] == s_dl[k] && v > best) best = v;
        }
        s_maxiou[k] = (best >= thr_min) ? 1.0f : 0.0f;
    }
    __syncthreads();
    // ordered compaction of the detections that can match at the loosest threshold
    {
        int base = 0;
        for (int k0 = 0; k0 < D; k0 += K2_THREADS) {
            int k = k0 + tid;
            bool f = (k < D) && s_maxiou[k] != 0.0f;
            unsigned m = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_warpcnt[wid] = __popc(m);
            __syncthreads();
            int off = base, tot = 0;
            for (int w = 0; w < K2_WARPS; ++w) { int v = s_warpcnt[w]; if (w < wid) off += v; tot += v; }
            __syncthreads();   // everyone has read s_maxiou / s_warpcnt before s_list (aliases s_kidx) is written
            if (f) s_list[off + __popc(m & ((1u << lane) - 1u))] = k;
            base += tot;
        }
        if (tid == 0) s_nlist = base;
        __syncthreads();
    }
    // default (unmatched) entries for every (a, t, d), zero padding beyond D
    {
        const size_t per_img = (size_t)BT_NUM_AREA * T * K;
        for (size_t q = tid; q < per_img; q += K2_THREADS) {
            int d = (int)(q % K);
            int a = (int)(q / ((size_t)T * K));
            uint8_t ig = 0;
            if (d < D) {
                double ar = s_db[d * 4 + 2] * s_db[d * 4 + 3];
                ig = (ar < area_lo[a] || ar > area_hi[a]) ? 1 : 0;
            }
            P.dt_match[(size_t)b * per_img + q] = 0;
            if (P.dt_ignore) P.dt_ignore[(size_t)b * per_img + q] = ig;
        }
    }
    __syncthreads();
    // greedy matching, one thread per (area range, threshold), over the compacted list only
    if (tid < BT_NUM_AREA * T) {
        const int a = tid / T, t = tid - a * T;
        const unsigned gign = s_gign[a];
        unsigned gm = 0;
        const double thr = fmin(P.thrs[t], 1 - 1e-10);
        const int nl = s_nlist;
        for (int li = 0; li < nl; ++li) {
            const int d = s_list[li];
            const int dl = s_dl[d];
            double best = thr;
            int m = -1;
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1 && m >= 0) break;          // matched a regular GT: stop at the first ignored one
                for (int g = 0; g < G; ++g) {
                    if ((((gign >> g) & 1u) != 0) != (pass == 1)) continue;
                    if ((gm >> g) & 1u) continue;
                    if (s_gl[g] != dl) continue;
                    double v = iou_in_smem ? s_iou[d * G + g] : bb_iou(s_db + d * 4, s_gb + g * 4);
                    if (v < best) continue;
                    best = v;
                    m = g;
                }
            }
            if (m >= 0) {
                gm |= 1u << m;
                size_t o = (((size_t)b * BT_NUM_AREA + a) * T + t) * K + d;
                P.dt_match[o] = m + 1;
                if (P.dt_ignore) P.dt_ignore[o] = (gign >> m) & 1u;
            }
        }
    }
}

size_t k2_smem_bytes(const BtParams &p, int smem_keys, int win) {
    size_t K = (size_t)p.max_det;
    return (size_t)smem_keys * 8 + (K + win) * sizeof(float4) + (K + win) * sizeof(float) + (2 * K + win) * sizeof(int);
}

int launch_nms_match(const BtParams &p, const BtIO &io, const Workspace &w, cudaStream_t s) {
    K2Params P{};
    P.N = p.num_anchors; P.nc = p.nc; P.nm = p.nm; P.C = 4 + p.nc + p.nm;
    P.cap = cand_capacity(&p); P.cap_pow2 = next_pow2(P.cap);
    P.max_det = p.max_det; P.max_gt = p.max_gt; P.class_mode = p.class_mode; P.layout = p.layout;
    float t = (float)p.iou_thres;
    if (!((double)t > p.iou_thres)) t = nextafterf(t, INFINITY);
    P.thr_up = t;
    P.early_out = p.iou_thres >= 0.0 ? 1 : 0;
    P.max_wh = p.max_wh;
    P.head = io.head; P.coeffs = io.coeffs;
    P.cand_box = w.cand_box; P.cand_score = w.cand_score; P.cand_label = w.cand_label;
    P.cand_anchor = w.cand_anchor; P.n_cand = io.n_cand; P.sort_keys = w.sort_keys;
    P.det_count = io.det_count; P.dets = io.dets; P.det_keep = io.det_keep;
    P.det_anchor = io.det_anchor; P.det_coeff = io.det_coeff;
    P.gt_count = io.gt_count; P.gt_boxes = io.gt_boxes; P.gt_labels = io.gt_labels;
    P.T = p.num_iou_thrs;
    for (int i = 0; i < p.num_iou_thrs; ++i) P.thrs[i] = p.iou_thrs[i];
    P.dt_match = io.dt_match; P.dt_ignore = io.dt_ignore; P.gt_ignore = io.gt_ignore;
    P.strip_done = w.strip_done; P.acc = w.acc; P.inst_area = io.inst_area; P.inst_inter = io.inst_inter;
    // shared key region: large enough for the sort of typical lists and for the COCO tables
    int need_coco = p.max_det * 4 + p.max_gt * 4 + p.max_det * 4;  // boxes + room for a [K x 4] IoU table at least
    int keys = next_pow2(P.cap) < SORT_SMEM_MAX ? next_pow2(P.cap) : SORT_SMEM_MAX;
    if (keys < need_coco) keys = need_coco;
    P.smem_keys = keys;
    // window of sorted candidates kept in shared memory: the whole list when it is small
    int win = (P.cap + NMS_CHUNK - 1) / NMS_CHUNK * NMS_CHUNK;
    if (win > 4096) win = 4096;
    P.win = win;
    P.centre_cull = (p.iou_thres >= 0.55) ? 1 : 0;
    P.crop = p.crop; P.PW = p.proto_w; P.PH = p.proto_h;
    P.rx = (float)((double)p.proto_w / (double)p.img_w);
    P.ry = (float)((double)p.proto_h / (double)p.img_h);
    P.det_region = w.det_region;
    size_t smem = k2_smem_bytes(p, keys, win);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(nms_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return BT_ERR_CUDA;
        attr_set = true;
    }
    if (smem > 200 * 1024) return BT_ERR_UNSUPPORTED;
    nms_match_kernel<<<p.batch, K2_THREADS, smem, s>>>(P);
    return cudaGetLastError() == cudaSuccess ? BT_OK : BT_ERR_CUDA;
}

}  // namespace bt
