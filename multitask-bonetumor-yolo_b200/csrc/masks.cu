// Kernel 3: mask assembly + segmentation counters in ONE pass over the prototypes.
//
// Reference statements replaced (paths under /root/reference/src):
//   M1 projector   running_main_v2.py:689-703 (Conv2d(32->1,k=1) -> F.interpolate bilinear x4,
//                  align_corners=False -> sigmoid -> >0.5 -> int), evaluate_model.py:160-171
//   M2 instance    test_model.py:80-85 (einsum coeff x protos -> bilinear -> sigmoid>0.5) with the
//                  Ultralytics process_mask crop at prototype resolution
//   counters       running_main_v2.py:704-713 (tp/fp/fn/tn, DiceScore) and test_model.py:15-23
//                  (per-image IoU / Dice with eps 1e-7)
//
// B200 mapping.  The prototypes are 2/3 of the path's compulsory HBM bytes and every other operand
// is tiny, so the kernel is a PERSISTENT streamer: one 1024-thread CTA per SM owns a contiguous
// range of strips (a strip = R "cell rows" of one image, a cell = the 4x4 output pixels between four
// neighbouring prototype pixels) and keeps the prototype rows it needs in a shared-memory RING of
// NS row slots (32 channels x PW floats each).  One warp feeds the ring with bulk async copies
// (TMA, cp.async.bulk + one mbarrier per slot) as soon as a strip has released its rows, so the
// loads always run a whole strip ahead of the arithmetic, every prototype row is read from HBM
// once (the row shared by two consecutive strips stays in the ring; r01d: the one-strip-per-CTA
// version sat 40 % of its time waiting for its own loads) and the GT-mask words of the next strip
// travel in registers meanwhile.
// Per strip the work is flattened over the CTA as two item lists found by prefix sums:
// (entry, row, 4-pixel group) for the K=32 contraction (sequential fp32 FMA, the oracle's
// summation order; TF32 tensor cores would break bit parity and the op is ~0.5 FLOP/B) and
// (entry, cell) for bilinear upsample + threshold.  An entry is a detection whose crop box
// touches the strip -- or the projector mask itself, which is just entry 0 with the projector
// weights, the bias as initial value and the whole strip as its box, so M1 costs no extra phases.
// Integer counters stay in registers across the strips of an image; output ownership per strip is
// exclusive, so there are no global atomics on pixels; the last CTA to finish an image turns the
// counters into Dice / IoU.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace bt {

constexpr int NM = 32;
constexpr int ECAP = 48;         // listed detections per round
constexpr int CF_PITCH = NM + 4; // coefficient row pitch in shared memory: 16-byte aligned rows (read as float4), rows 4 banks apart
constexpr int NS_MAX = 12;       // ring slots

struct K3Params {
    int B, S_h, S_w, PH, PW, R, NS, K, crop, gt_f32, nstrips, total_strips, scr_cap;
    float bias;
    const float *protos, *proj_weight, *det_coeff;
    const int32_t *det_count;
    const short4 *det_region;
    const void *masks_gt;
    int32_t *strip_done, *acc, *inst_area, *inst_inter;
    long long *seg_cnt4, *uni_cnt4, *seg_img3, *uni_img3;
    float *seg_dice, *seg_iou, *uni_dice, *uni_iou;
    uint8_t *seg_mask, *uni_mask;
    float *seg_logits;
    double *seg_prob_sum;   // [B] optional: sum of sigmoid(logit) over the projector mask's foreground pixels
    // shared-memory offsets (bytes)
    int off_lm, off_scr, off_gtrow, off_gtc, off_unc, off_list, off_cf, wpr;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// One prototype row of all 32 channels ([NM][PW] box of the [B*NM, PH, PW] tensor) -> one ring slot.
__device__ __forceinline__ void tma_row_g2s(void *dst, const CUtensorMap *tm, int row, int chan0, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(chan0), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// torch's bilinear kernel order, pinned in the oracle: fma(h0, fma(w0,v00,w1*v01), h1*fma(w0,v10,w1*v11))
__device__ __forceinline__ float lerp_row(float a, float b, float w0, float w1) {
    return __fmaf_rn(w0, a, __fmul_rn(w1, b));
}

// Evaluate one cell: 4 corner values -> 4x4 thresholded output pixels, bit (ry*4+rx).
// Border cells (index -1) use interpolation weight 0 (source coordinate clamped to 0).
template <bool LOG>
__device__ __forceinline__ unsigned cell_bits(float v00, float v01, float v10, float v11, bool border_y, bool border_x,
                                              float (&logits)[16]) {
    float top[4], bot[4];
#pragma unroll
    for (int rx = 0; rx < 4; ++rx) {
        float w1 = border_x ? 0.0f : (0.125f + 0.25f * rx);
        float w0 = 1.0f - w1;
        top[rx] = lerp_row(v00, v01, w0, w1);
        bot[rx] = lerp_row(v10, v11, w0, w1);
    }
    unsigned bits = 0;
#pragma unroll
    for (int ry = 0; ry < 4; ++ry) {
        float h1 = border_y ? 0.0f : (0.125f + 0.25f * ry);
        float h0 = 1.0f - h1;
#pragma unroll
        for (int rx = 0; rx < 4; ++rx) {
            float v = __fmaf_rn(h0, top[rx], __fmul_rn(h1, bot[rx]));
            if (LOG) logits[ry * 4 + rx] = v;
            if (sigmoid_gt_half(v)) bits |= 1u << (ry * 4 + rx);
        }
    }
    return bits;
}

// Valid-pixel mask of a cell: border cells (-1) own output rows/cols {0,1}; the last cell row/col
// owns only the two pixels left before the image edge.
__device__ __forceinline__ unsigned cell_valid(int ci, int cj, int S_h, int S_w) {
    const int nry = (ci < 0) ? 2 : min(4, S_h - (4 * ci + 2));
    const int nrx = (cj < 0) ? 2 : min(4, S_w - (4 * cj + 2));
    const unsigned rowm = (1u << nrx) - 1u;
    unsigned m = 0;
    for (int r = 0; r < nry; ++r) m |= rowm << (4 * r);
    return m;
}

__device__ __forceinline__ uint32_t pack_u8(const uint4 &v, int half) {
    uint32_t wv[4] = {v.x, v.y, v.z, v.w}, bits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t x = wv[j];   // byte != 0 -> bit: fold each byte to its low bit, then gather the four low bits
        x = (x | (x >> 4)) & 0x0f0f0f0fu;
        x = (x | (x >> 2)) & 0x03030303u;
        x = (x | (x >> 1)) & 0x01010101u;
        bits |= ((x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xfu) << (16 * half + 4 * j);
    }
    return bits;
}

// Strip geometry (s = strip index inside its image).
struct StripGeo {
    int ci_lo, ci_hi, ncr_all, p_lo, p_hi, nrows, y_lo, nyrows;
};
__device__ __forceinline__ StripGeo strip_geo(int s, int R, int PH, int S_h) {
    StripGeo g;
    g.ci_lo = (s == 0) ? -1 : s * R;
    g.ci_hi = min(s * R + R - 1, PH - 1);
    g.ncr_all = g.ci_hi - g.ci_lo + 1;
    g.p_lo = s * R;
    g.p_hi = min(g.ci_hi + 1, PH - 1);
    g.nrows = g.p_hi - g.p_lo + 1;
    g.y_lo = (s == 0) ? 0 : 4 * s * R + 2;
    const int y_hi = (g.ci_hi == PH - 1) ? S_h : 4 * (g.ci_hi + 1) + 2;
    g.nyrows = y_hi - g.y_lo;
    return g;
}

// One (detection, row, 4-pixel group) item of the K=32 contraction: logits of 4 prototype pixels,
// zero outside the crop box.
template <int PW4>
__device__ __forceinline__ float4 contract4(const float4 *pp, const float *cf, int pw4_rt) {
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const int st = PW4 > 0 ? PW4 : pw4_rt;
    const float4 *cf4 = reinterpret_cast<const float4 *>(cf);
#pragma unroll
    for (int i4 = 0; i4 < NM / 4; ++i4) {
        const float4 w4 = cf4[i4];
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float4 v = pp[(i4 * 4 + u) * st];
            const float w = wv[u];
            acc.x = __fmaf_rn(w, v.x, acc.x); acc.y = __fmaf_rn(w, v.y, acc.y);
            acc.z = __fmaf_rn(w, v.z, acc.z); acc.w = __fmaf_rn(w, v.w, acc.w);
        }
    }
    return acc;
}

__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Piece tables of a round (= up to ECAP listed detections of a strip).
struct Tab {
    short4 ereg[ECAP];
    int pxoff[ECAP + 1], celloff[ECAP + 1];
    short rlo[ECAP], rhi[ECAP], clo[ECAP], chi[ECAP];
    short pra[ECAP], pa[ECAP], npc[ECAP], cia[ECAP], ncc[ECAP];
    float inpc[ECAP], incc[ECAP];
    int area[ECAP], inter[ECAP];
    int nb, be[ECAP + 1];   // batches of a round: entries [be[i], be[i+1])
    int nlist;
};
// inside the kernel `T` is the table set in use (a reference, or a lambda parameter of that name)
#define s_ereg T.ereg
#define s_pxoff T.pxoff
#define s_celloff T.celloff
#define s_rlo T.rlo
#define s_rhi T.rhi
#define s_clo T.clo
#define s_chi T.chi
#define s_pra T.pra
#define s_pa T.pa
#define s_npc T.npc
#define s_cia T.cia
#define s_ncc T.ncc
#define s_inpc T.inpc
#define s_incc T.incc
#define s_area T.area
#define s_inter T.inter
#define s_nb T.nb
#define s_be T.be
#define s_nlist T.nlist

// TPW > 0: compile-time prototype width (shared-memory strides become immediates); 0: run-time.
template <int TPW, int K3_THREADS, int MINB>
__global__ void __launch_bounds__(K3_THREADS, MINB)
masks_kernel(const __grid_constant__ K3Params P, const __grid_constant__ CUtensorMap tmap) {
    constexpr int K3_WARPS = K3_THREADS / 32;
    constexpr int RPL = 10;   // crop regions cached per lane of warp 0 (detections lane, lane + 32, ...)
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // TMA destinations must be 128-byte aligned: align the dynamic region by hand (128 spare bytes are allocated)
    unsigned char *smem = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    __shared__ __align__(8) uint64_t s_bar[NS_MAX];
    __shared__ float s_w[NM];
    __shared__ int s_red[K3_WARPS][5];
    __shared__ int s_last;
    __shared__ int s_rowoff[16];   // float offset of the strip's prototype rows inside the ring
    // per-round piece tables, double buffered: warp 0 builds the next strip's set during this strip's contraction
    __shared__ Tab s_tab[2];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int PW = TPW > 0 ? TPW : P.PW, R = P.R, NS = P.NS;
    const int PH = P.PH, S_w = P.S_w, S_h = P.S_h, K = P.K;
    const int rowsmax = R + 1;
    const int ncc_all = PW + 1;    // cells per cell row (incl. the border column -1)
    const int SLOT = NM * PW;      // floats per ring slot
    const int SCR_GRP = P.scr_cap >> 2;

    float *s_ring = reinterpret_cast<float *>(smem);                           // [NS][NM][PW]
    float *s_lm = reinterpret_cast<float *>(smem + P.off_lm);                  // [R+1][PW] projector logits
    float *s_scr = reinterpret_cast<float *>(smem + P.off_scr);                // [scr_cap + (R+1)*PW]
    uint32_t *s_gtrow = reinterpret_cast<uint32_t *>(smem + P.off_gtrow);      // [4R+2][wpr+1] row bits (bit x)
    uint32_t *s_gtc = reinterpret_cast<uint32_t *>(smem + P.off_gtc);          // [R+1][PW+1] cell bits
    uint32_t *s_unc = reinterpret_cast<uint32_t *>(smem + P.off_unc);          // [R+1][PW+1]
    unsigned short *s_list2 = reinterpret_cast<unsigned short *>(smem + P.off_list);  // [2][KP] detections listed on a strip
    const int KP = (K + 7) & ~7;
    float *s_cf = reinterpret_cast<float *>(smem + P.off_cf);                  // [ECAP][CF_PITCH]
    const int wpr = P.wpr, tp = wpr + 1;

    // ---- this CTA's contiguous range of strips
    const int g0 = (int)(((long long)blockIdx.x * P.total_strips) / gridDim.x);
    const int g1 = (int)(((long long)(blockIdx.x + 1) * P.total_strips) / gridDim.x);
    if (g0 >= g1) return;

    BT_PHASE_INIT();
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- producer state (thread 0).  Rows are loaded in the order the strips consume them; the row
    // shared by two consecutive strips of an image is loaded once.  Row q of this sequence lives in
    // slot q % NS.
    int pg = g0, pb = g0 / P.nstrips, ps = g0 - (g0 / P.nstrips) * P.nstrips, prow = ps * R, pseq = 0;
    int seq_base = 0;   // sequence number of the current strip's first row
    auto top_up = [&]() {
        // fill every free slot (rows before seq_base are released), one TMA per row
        while (pseq < seq_base + NS && pg < g1) {
            const int p_hi = min(min(ps * R + R - 1, PH - 1) + 1, PH - 1);
            const int slot = pseq % NS;
            mbar_expect_tx(&s_bar[slot], (uint32_t)(SLOT * sizeof(float)));
            tma_row_g2s(s_ring + (size_t)slot * SLOT, &tmap, prow, pb * NM, &s_bar[slot]);
            ++pseq;
            if (prow < p_hi) {
                ++prow;
            } else {
                // next strip that brings new rows (a last strip of one row only re-uses its predecessor's)
                for (;;) {
                    ++pg;
                    if (++ps == P.nstrips) { ps = 0; ++pb; }
                    if (pg >= g1) break;
                    const int start = (ps == 0) ? 0 : ps * R + 1;   // same image: the strip's first row is already in the ring
                    if (start <= min(min(ps * R + R - 1, PH - 1) + 1, PH - 1)) { prow = start; break; }
                }
            }
        }
    };
    if (tid == 0) top_up();
    if (tid < NM) s_w[tid] = __ldg(P.proj_weight + tid);

    // ---- L2 prefetch of what the next strip will read: its GT words (32 pixels per thread) and its new
    // prototype rows (their TMA is only issued half-way through this strip: a register prefetch of the GT
    // words was spilled to local memory by ptxas and stalled on the load, r02c)
    auto prefetch_next = [&](int g, int b, int s) {
        if (g >= g1) return;
        const StripGeo G = strip_geo(s, R, PH, S_h);
        if (!P.gt_f32 && tid < G.nyrows * wpr && (tid & 3) == 0) {   // one 128-byte line per 4 threads
            const int yr = tid / wpr, w = tid - yr * wpr;
            const uint8_t *gp = static_cast<const uint8_t *>(P.masks_gt) + ((size_t)b * S_h + (G.y_lo + yr)) * S_w + (size_t)w * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gp));
        }
        // rows p_lo+1 .. p_hi of the next strip (p_lo is shared with this one), NM channels, PW*4 bytes each
        const int lines_per_row = (PW * 4 + 127) >> 7, nnew = G.p_hi - G.p_lo;
        for (int q = tid; q < nnew * NM * lines_per_row; q += K3_THREADS) {
            const int ln = q % lines_per_row, ch = (q / lines_per_row) % NM, rr = q / (lines_per_row * NM);
            const float *pp = P.protos + (((size_t)b * NM + ch) * PH + G.p_lo + 1 + rr) * PW + ln * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
        }
    };
    int b = g0 / P.nstrips, s = g0 - b * P.nstrips;

    double psum = 0.0;             // this thread's share of seg_prob_sum of the current image
    int c5[5] = {0, 0, 0, 0, 0};   // seg inter, seg P, G, uni inter, uni P of the current image (this thread's share)
    int cur_b = -1, strips_of_b = 0;
    bool prebuilt = false;         // the current strip's list + tables were built during the previous strip
    short4 myreg[RPL];             // warp 0: crop regions of detections lane + 32 i of the current image

    // counters of image `b` -> global accumulators; the CTA that completes the image finalises it
    auto flush_image = [&](int b, int nstrips_done) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            int v = c5[i];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane == 0) s_red[wid][i] = v;
            c5[i] = 0;
        }
        if (P.seg_prob_sum) {
            double v = psum;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
            if (lane == 0 && v != 0.0) atomicAdd(&P.seg_prob_sum[b], v);
            psum = 0.0;
        }
        __syncthreads();
        if (tid < 5) {
            int v = 0;
#pragma unroll
            for (int w = 0; w < K3_WARPS; ++w) v += s_red[w][tid];
            if (v) atomicAdd(&P.acc[b * 8 + tid], v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&P.strip_done[b], nstrips_done) + nstrips_done == P.nstrips);
        __syncthreads();
        if (s_last && tid < 2) {
            // (test_model.py:15-23)
            __threadfence();
            long long inter = atomicAdd(&P.acc[b * 8 + 3 * tid], 0), pp = atomicAdd(&P.acc[b * 8 + 3 * tid + 1], 0);
            long long gg = atomicAdd(&P.acc[b * 8 + 2], 0);   // |G| is shared by both masks
            long long total = (long long)S_h * S_w;
            long long *img3 = tid ? P.uni_img3 : P.seg_img3;
            long long *cnt4 = tid ? P.uni_cnt4 : P.seg_cnt4;
            float *dice = tid ? P.uni_dice : P.seg_dice, *iou = tid ? P.uni_iou : P.seg_iou;
            if (img3) { img3[b * 3 + 0] = inter; img3[b * 3 + 1] = pp; img3[b * 3 + 2] = gg; }
            if (cnt4) {
                atomicAdd((unsigned long long *)&cnt4[0], (unsigned long long)inter);
                atomicAdd((unsigned long long *)&cnt4[1], (unsigned long long)(pp - inter));
                atomicAdd((unsigned long long *)&cnt4[2], (unsigned long long)(gg - inter));
                atomicAdd((unsigned long long *)&cnt4[3], (unsigned long long)(total - pp - gg + inter));
            }
            const float fi = (float)inter, fu = (float)(pp + gg - inter);
            if (iou) iou[b] = __fdiv_rn(__fadd_rn(fi, 1e-7f), __fadd_rn(fu, 1e-7f));
            if (dice) dice[b] = __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, fi), 1e-7f), __fadd_rn(__fadd_rn((float)pp, (float)gg), 1e-7f));
        }
        __syncthreads();
    };

    for (int g = g0; g < g1; ++g) {
        Tab &T = s_tab[g & 1], &TN = s_tab[(g + 1) & 1];
        unsigned short *s_list = s_list2 + (g & 1) * KP, *s_list_n = s_list2 + ((g + 1) & 1) * KP;
        const StripGeo G = strip_geo(s, R, PH, S_h);
        const int nb = (s + 1 == P.nstrips) ? b + 1 : b, ns = (s + 1 == P.nstrips) ? 0 : s + 1;   // next strip
        const int ci_lo = G.ci_lo, ci_hi = G.ci_hi, ncr_all = G.ncr_all, p_lo = G.p_lo, nrows = G.nrows;
        const int y_lo = G.y_lo, nyrows = G.nyrows;
        const int next_base = seq_base + ((g + 1 < g1 && nb == b) ? nrows - 1 : nrows);   // rows shift (same image keeps the last row)
        const bool prebuild_next = (g + 1 < g1) && nb == b;
        const StripGeo GN = strip_geo(ns, R, PH, S_h);

        // ---- (1) new image: flush the previous one, warp 0 fetches this image's crop regions
        if (b != cur_b) {
            if (cur_b >= 0) flush_image(cur_b, strips_of_b);
            cur_b = b; strips_of_b = 0;
            if (wid == 0) {
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    const int k = i * 32 + lane;
                    myreg[i] = (k < K) ? __ldg(P.det_region + (size_t)b * K + k) : make_short4(1, 0, 1, 0);
                }
            }
        }
        ++strips_of_b;
        // GT words (L2-resident: prefetched during the previous strip) -> row bits, union tile cleared
        if (tid < rowsmax) s_rowoff[tid] = ((seq_base + tid) % NS) * SLOT;
        for (int i = tid; i < rowsmax * ncc_all; i += K3_THREADS) s_unc[i] = 0;
        for (int q = tid; q < nyrows; q += K3_THREADS) s_gtrow[q * tp + wpr] = 0;   // pad word
        if (!P.gt_f32) {
            for (int q = tid; q < nyrows * wpr; q += K3_THREADS) {
                const int yr = q / wpr, w = q - yr * wpr;
                const uint4 *gp = reinterpret_cast<const uint4 *>(static_cast<const uint8_t *>(P.masks_gt) +
                                                                  ((size_t)b * S_h + (y_lo + yr)) * S_w + (size_t)w * 32);
                s_gtrow[yr * tp + w] = pack_u8(__ldg(gp), 0) | pack_u8(__ldg(gp + 1), 1);
            }
        } else {
            for (int q = tid; q < nyrows * wpr; q += K3_THREADS) {
                const int yr = q / wpr, w = q - yr * wpr;
                const float4 *gp = reinterpret_cast<const float4 *>(static_cast<const float *>(P.masks_gt) +
                                                                    ((size_t)b * S_h + (y_lo + yr)) * S_w + (size_t)w * 32);
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 v = __ldg(gp + i);
                    bits |= ((int)v.x != 0 ? 1u : 0u) << (4 * i) | ((int)v.y != 0 ? 1u : 0u) << (4 * i + 1) |
                            ((int)v.z != 0 ? 1u : 0u) << (4 * i + 2) | ((int)v.w != 0 ? 1u : 0u) << (4 * i + 3);
                }
                s_gtrow[yr * tp + w] = bits;
            }
        }
        prefetch_next(g + 1, nb, ns);   // next strip's GT words and prototype rows -> L2
        __syncthreads();
        BT_PHASE_MARK(2, 0);   // GT words, tiles

        // piece tables of the entries [r0, r0 + nch) of the strip's list (warp 0)
        auto build_tables = [&](Tab &T, const unsigned short *s_list, int r0, int nch, bool from_list, int ci_lo, int ci_hi) {
            int npx[2] = {0, 0}, ncell[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                if (e < nch) {
                    short4 rg;
                    if (from_list) { rg = __ldg(P.det_region + (size_t)b * K + s_list[r0 + e]); s_ereg[e] = rg; }
                    else rg = s_ereg[e];
                    const int r_lo = rg.x, r_hi = rg.y, c_lo = rg.z, c_hi = rg.w;
                    const int ci_a = max(r_lo - 1, ci_lo), ci_b = min(r_hi, ci_hi);
                    const int pr_a = max(ci_a, 0), pr_b = min(ci_b + 1, PH - 1);
                    const int ja = c_lo - 1;
                    const int pa = max(ja, 0), pb2 = min(c_hi + 1, PW - 1);
                    // scratch rows are stored in aligned groups of 4 prototype columns
                    const int ga = pa >> 2, ngrp = (pb2 >> 2) - ga + 1;
                    const int npr = pr_b - pr_a + 1;
                    const int ncr = ci_b - ci_a + 1, ncc = c_hi - ja + 1;
                    s_rlo[e] = r_lo; s_rhi[e] = r_hi; s_clo[e] = c_lo; s_chi[e] = c_hi;
                    s_pra[e] = pr_a; s_pa[e] = 4 * ga; s_npc[e] = 4 * ngrp; s_cia[e] = ci_a; s_ncc[e] = ncc;
                    const int nst = (ncc + 3) >> 2;   // cell items are runs of 4 cells of a cell row
                    s_inpc[e] = 1.0f / (float)ngrp; s_incc[e] = 1.0f / (float)nst;
                    s_area[e] = 0; s_inter[e] = 0;
                    npx[h] = npr * ngrp; ncell[h] = ncr * nst;
                }
            }
            // inclusive scans over the (up to) 64 entries: lanes, then the second half on top of the first
            int ipx[2] = {npx[0], npx[1]}, icl[2] = {ncell[0], ncell[1]};
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int v = __shfl_up_sync(0xffffffffu, ipx[h], d), u = __shfl_up_sync(0xffffffffu, icl[h], d);
                    if (lane >= d) { ipx[h] += v; icl[h] += u; }
                }
            }
            const int tpx = __shfl_sync(0xffffffffu, ipx[0], 31), tcl = __shfl_sync(0xffffffffu, icl[0], 31);
            ipx[1] += tpx; icl[1] += tcl;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = lane + 32 * h;
                if (e < ECAP) { s_pxoff[e + 1] = ipx[h]; s_celloff[e + 1] = icl[h]; }
            }
            if (lane == 0) { s_pxoff[0] = 0; s_celloff[0] = 0; }
            // batches: consecutive entries whose first scratch group falls into the same SCR_GRP window
            unsigned first[2];
            {
                const int bid0 = (ipx[0] - npx[0]) / SCR_GRP, bid1 = (ipx[1] - npx[1]) / SCR_GRP;
                const int pb0 = __shfl_up_sync(0xffffffffu, bid0, 1), last0 = __shfl_sync(0xffffffffu, bid0, 31);
                int pb1 = __shfl_up_sync(0xffffffffu, bid1, 1);
                if (lane == 0) pb1 = last0;
                first[0] = __ballot_sync(0xffffffffu, lane < nch && (lane == 0 || bid0 != pb0));
                first[1] = __ballot_sync(0xffffffffu, lane + 32 < nch && bid1 != pb1);
            }
            if (lane == 0) {
                unsigned long long m = ((unsigned long long)first[1] << 32) | first[0];
                int nbt = 0;
                while (m) { s_be[nbt++] = __ffsll((long long)m) - 1; m &= m - 1ull; }
                s_be[nbt] = nch;
                s_nb = nbt;
            }
        };

        // detection list of a strip of this image (ordered ballot compaction of the cached regions) and the
        // first round's piece tables (warp 0)
        auto list_and_tables = [&](Tab &T, unsigned short *s_list, int lo, int hi) {
            int n = 0;
#pragma unroll
            for (int i = 0; i < RPL; ++i) {
                const short4 rg = myreg[i];
                const bool ok = rg.x <= rg.y && rg.z <= rg.w && max((int)rg.x - 1, lo) <= min((int)rg.y, hi);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int pos = n + __popc(m & ((1u << lane) - 1u));
                    s_list[pos] = (unsigned short)(i * 32 + lane);
                    if (pos < ECAP) s_ereg[pos] = rg;
                }
                n += __popc(m);
            }
            for (int k0 = RPL * 32; k0 < K; k0 += 32) {   // more detections than the register cache holds
                const int k = k0 + lane;
                const short4 rg = (k < K) ? __ldg(P.det_region + (size_t)b * K + k) : make_short4(1, 0, 1, 0);
                const bool ok = rg.x <= rg.y && rg.z <= rg.w && max((int)rg.x - 1, lo) <= min((int)rg.y, hi);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const int pos = n + __popc(m & ((1u << lane) - 1u));
                    s_list[pos] = (unsigned short)k;
                    if (pos < ECAP) s_ereg[pos] = rg;
                }
                n += __popc(m);
            }
            if (lane == 0) s_nlist = n;
            __syncwarp();
            build_tables(T, s_list, 0, min(n, ECAP), false, lo, hi);
        };
        // GT cells from the row bits: bit x of output row y -> bit ry*4+rx of cell (ci, cj)
        auto gt_cells = [&]() {
            for (int q = tid; q < ncr_all * ncc_all; q += K3_THREADS) {
                const int cr = q / ncc_all, cj = q - cr * ncc_all - 1, ci = ci_lo + cr;
                const int ybase = (ci < 0) ? 0 : 4 * ci + 2, xbase = (cj < 0) ? 0 : 4 * cj + 2;
                const int nry = (ci < 0) ? 2 : min(4, S_h - ybase), nrx = (cj < 0) ? 2 : min(4, S_w - xbase);
                unsigned bits = 0;
                for (int ry = 0; ry < nry; ++ry) {
                    const uint32_t *row = s_gtrow + (ybase + ry - y_lo) * tp + (xbase >> 5);
                    unsigned gb = __funnelshift_r(row[0], row[1], xbase & 31) & ((1u << nrx) - 1u);
                    bits |= gb << (4 * ry);
                }
                s_gtc[q] = bits;
            }
        };
        // ---- (2) only at the first strip of an image in this CTA: warp 0 builds the strip's list + tables
        // (every other strip's were built by warp 0 during the previous strip's contraction phase)
        if (!prebuilt) {
            if (wid == 0) list_and_tables(T, s_list, ci_lo, ci_hi);
            __syncthreads();
        }
        BT_PHASE_MARK(2, 1);   // list + tables (first strip of an image)
        const int nent = s_nlist;
        const int nst_all = (ncc_all + 3) >> 2;     // runs of 4 cells per cell row of the projector mask
        const int nM1 = ncr_all * nst_all;

        for (int r0 = 0; r0 == 0 || r0 < nent; r0 += ECAP) {
            const int nch = min(ECAP, nent - r0);
            if (r0 > 0) {
                __syncthreads();   // the previous round's tables are still being read
                if (wid == 0) build_tables(T, s_list, r0, nch, true, ci_lo, ci_hi);
                __syncthreads();
            }
            // coefficients of the round: asynchronous 4-byte copies into the pitched table
            for (int q = tid; q < nch * NM; q += K3_THREADS) {
                const int e = q >> 5, i = q & 31;
                cp_async4(s_cf + e * CF_PITCH + i, P.det_coeff + ((size_t)b * K + s_list[r0 + e]) * NM + i);
            }
            if (r0 == 0) {
                gt_cells();
                // ---- (3) the strip's prototype rows (normally long since landed), then the M1 projection:
                // bias + sum_k w_k p_k, sequential fmaf (== torch conv2d, pinned).  Four neighbouring
                // pixels per thread: one 16-byte shared load per channel feeds four FMA chains.
                for (int i = 0; i < nrows; ++i) mbar_wait(&s_bar[(seq_base + i) % NS], (uint32_t)(((seq_base + i) / NS) & 1));
                for (int q = tid; q < (nrows * PW) >> 2; q += K3_THREADS) {
                    const int rr = q / (PW >> 2), cg = q - rr * (PW >> 2);
                    float4 acc = make_float4(P.bias, P.bias, P.bias, P.bias);
                    const float4 *pp = reinterpret_cast<const float4 *>(s_ring + s_rowoff[rr]) + cg;
#pragma unroll
                    for (int k = 0; k < NM; ++k) {
                        const float4 v = pp[k * (PW >> 2)];
                        const float w = s_w[k];
                        acc.x = __fmaf_rn(w, v.x, acc.x); acc.y = __fmaf_rn(w, v.y, acc.y);
                        acc.z = __fmaf_rn(w, v.z, acc.z); acc.w = __fmaf_rn(w, v.w, acc.w);
                    }
                    reinterpret_cast<float4 *>(s_lm)[q] = acc;
                }
            }
            cp_async_wait_all();
            __syncthreads();
            BT_PHASE_MARK(2, 2);   // coefficients + M1 projection
            const int nbatch = (nch > 0) ? s_nb : 0;
            for (int bi = 0; bi == 0 || bi < nbatch; ++bi) {
                const int e0 = (nch > 0) ? s_be[bi] : 0, e1 = (nch > 0) ? s_be[bi + 1] : 0;
                const int px0 = s_pxoff[e0], npx = s_pxoff[e1] - px0;
                // ---- (4) one phase: M1 cells (first pass only) and the contraction items of the batch
                const int nm1 = (r0 == 0 && bi == 0) ? nM1 : 0;
                // warp 0 spends the strip's first contraction phase on the next strip's detection list + tables
                const bool w0_builds = prebuild_next && r0 == 0 && bi == 0;
                if (w0_builds && wid == 0) list_and_tables(TN, s_list_n, GN.ci_lo, GN.ci_hi);
                for (int q = w0_builds ? tid - 32 : tid; q < nm1 + npx; q += w0_builds ? K3_THREADS - 32 : K3_THREADS) {
                    if (q < 0) break;   // warp 0 (building the tables)
                    if (q < nm1) {
                        // projector mask: run of 4 cells of a cell row (exclusive owner of its cell tile, + optional logits)
                        const int cr = q / nst_all, cj0 = 4 * (q - cr * nst_all) - 1, ci = ci_lo + cr;
                        const int nk = min(4, PW - cj0);
                        const int rr0 = max(ci, 0) - p_lo, rr1 = ((ci < 0) ? 1 : min(ci + 1, PH - 1)) - p_lo;
                        const int cb = max(cj0, 0), sh = (cj0 < 0) ? 1 : 0;
                        float v0[5], v1[5];
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const int c = min(cb + i, PW - 1);
                            v0[i] = s_lm[rr0 * PW + c]; v1[i] = s_lm[rr1 * PW + c];
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k >= nk) break;
                            const int cj = cj0 + k;
                            const float a0 = sh ? v0[k ? k - 1 : 0] : v0[k], a1 = sh ? v1[k ? k - 1 : 0] : v1[k];
                            float b0 = sh ? v0[k ? k : 1] : v0[k + 1], b1 = sh ? v1[k ? k : 1] : v1[k + 1];
                            if (cj >= PW - 1) { b0 = a0; b1 = a1; }
                            float lg[16];
                            unsigned bits = (P.seg_logits || P.seg_prob_sum) ? cell_bits<true>(a0, b0, a1, b1, ci < 0, cj < 0, lg)
                                                                             : cell_bits<false>(a0, b0, a1, b1, ci < 0, cj < 0, lg);
                            if (ci < 0 || cj < 0 || ci == PH - 1 || cj == PW - 1) bits &= cell_valid(ci, cj, S_h, S_w);
                            // the projector mask needs no tile: its counters are taken here (GT cells are complete)
                            c5[0] += __popc(bits & s_gtc[cr * ncc_all + cj + 1]);
                            c5[1] += __popc(bits);
                            if (P.seg_prob_sum && bits) {
                                float ps = 0.0f;   // v3 seg-mAP score numerator: sigmoid over the foreground pixels of the cell
#pragma unroll
                                for (int k = 0; k < 16; ++k)
                                    if ((bits >> k) & 1u) ps += 1.0f / (1.0f + __expf(-lg[k]));
                                psum += (double)ps;
                            }
                            if (P.seg_logits || P.seg_mask) {
                                const int ybase = (ci < 0) ? 0 : 4 * ci + 2, xbase = (cj < 0) ? 0 : 4 * cj + 2;
                                const int nry = (ci < 0) ? 2 : min(4, S_h - ybase), nrx = (cj < 0) ? 2 : min(4, S_w - xbase);
                                for (int ry = 0; ry < nry; ++ry) {
                                    const size_t o = ((size_t)b * S_h + ybase + ry) * S_w + xbase;
                                    for (int rx = 0; rx < nrx; ++rx) {
                                        if (P.seg_logits) P.seg_logits[o + rx] = lg[ry * 4 + rx];
                                        if (P.seg_mask) P.seg_mask[o + rx] = (bits >> (ry * 4 + rx)) & 1u;
                                    }
                                }
                            }
                        }
                        continue;
                    }
                    const int qq = q - nm1;
                    int lo = e0, hi = e1;   // last entry with pxoff <= px0 + qq
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_pxoff[mid] - px0 <= qq) lo = mid; else hi = mid; }
                    const int e = lo, loc = qq - (s_pxoff[e] - px0);
                    const int ngrp = s_npc[e] >> 2;
                    const int rr = __float2int_rz(((float)loc + 0.5f) * s_inpc[e]), gg = loc - rr * ngrp;
                    const int r = s_pra[e] + rr, c = s_pa[e] + 4 * gg;
                    const int c_lo = s_clo[e], c_hi = s_chi[e];
                    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (r >= s_rlo[e] && r <= s_rhi[e] && c + 3 >= c_lo && c <= c_hi) {
                        acc = contract4<(TPW >> 2)>(reinterpret_cast<const float4 *>(s_ring + s_rowoff[r - p_lo] + c),
                                                    s_cf + e * CF_PITCH, PW >> 2);
                        if (c < c_lo || c > c_hi) acc.x = 0.0f;
                        if (c + 1 < c_lo || c + 1 > c_hi) acc.y = 0.0f;
                        if (c + 2 < c_lo || c + 2 > c_hi) acc.z = 0.0f;
                        if (c + 3 < c_lo || c + 3 > c_hi) acc.w = 0.0f;
                    }
                    reinterpret_cast<float4 *>(s_scr)[qq] = acc;
                }
                __syncthreads();
                BT_PHASE_MARK(2, 9);   // M1 cells + contraction
                if (r0 + ECAP >= nent && bi + 1 >= nbatch && tid == 0) {
                    // last contraction of the strip: its rows are released (the cells only read the scratch),
                    // the ring is topped up while the strip finishes
                    seq_base = next_base;
                    top_up();
                }
                // ---- (5) upsample + threshold: items = (detection, cell row, run of 4 cells); the search, the
                // table reads and the two rows of corner logits are shared by the four cells
                const int cl0 = s_celloff[e0], ncell = s_celloff[e1] - cl0;
                for (int q = tid; q < ncell; q += K3_THREADS) {
                    int lo = e0, hi = e1;
                    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_celloff[mid] - cl0 <= q) lo = mid; else hi = mid; }
                    const int e = lo, loc = q - (s_celloff[e] - cl0);
                    const int ncc = s_ncc[e], npc = s_npc[e], nst = (ncc + 3) >> 2;
                    const int cr = __float2int_rz(((float)loc + 0.5f) * s_incc[e]);
                    const int ci = s_cia[e] + cr, cj0 = s_clo[e] - 1 + 4 * (loc - cr * nst);
                    const int nk = min(4, s_clo[e] - 1 + ncc - cj0);
                    const int pr_a = s_pra[e], pa = s_pa[e];
                    const float *scr = s_scr + 4 * (s_pxoff[e] - px0);
                    const int rr0 = max(ci, 0) - pr_a, rr1 = ((ci < 0) ? 1 : min(ci + 1, PH - 1)) - pr_a;
                    const int cb = max(cj0, 0) - pa, sh = (cj0 < 0) ? 1 : 0;
                    float v0[5], v1[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const int c = min(cb + i, npc - 1);
                        v0[i] = scr[rr0 * npc + c]; v1[i] = scr[rr1 * npc + c];
                    }
                    int area = 0, inter = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k >= nk) break;
                        const int cj = cj0 + k;
                        // corner columns of cell k: (k, k+1) from the strip's base; the border cell (cj = -1) reads the
                        // same columns (0, 1) as cell 0; in the last column the right corner is the left one
                        const float a0 = sh ? v0[k ? k - 1 : 0] : v0[k], a1 = sh ? v1[k ? k - 1 : 0] : v1[k];
                        float b0 = sh ? v0[k ? k : 1] : v0[k + 1], b1 = sh ? v1[k ? k : 1] : v1[k + 1];
                        if (cj >= PW - 1) { b0 = a0; b1 = a1; }
                        float lg[16];
                        unsigned bits = cell_bits<false>(a0, b0, a1, b1, ci < 0, cj < 0, lg);
                        if (bits == 0) continue;
                        if (ci < 0 || cj < 0 || ci == PH - 1 || cj == PW - 1) bits &= cell_valid(ci, cj, S_h, S_w);
                        if (bits == 0) continue;
                        const int cell = (ci - ci_lo) * ncc_all + cj + 1;
                        atomicOr(&s_unc[cell], bits);
                        area += __popc(bits);
                        inter += __popc(bits & s_gtc[cell]);
                    }
                    if (area) atomicAdd(&s_area[e], area);
                    if (inter) atomicAdd(&s_inter[e], inter);
                }
                __syncthreads();
                BT_PHASE_MARK(2, 10);  // cells
            }
            if (wid == 0) {   // warp 0 alone reads the round's tables here
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = lane + 32 * h;
                    if (e < nch) {
                        const int k = s_list[r0 + e];
                        if (s_area[e] && P.inst_area) atomicAdd(&P.inst_area[(size_t)b * K + k], s_area[e]);
                        if (s_inter[e] && P.inst_inter) atomicAdd(&P.inst_inter[(size_t)b * K + k], s_inter[e]);
                    }
                }
            }
        }

        // ---- (6) integer counters of the strip (kept in registers) + optional dense mask output
        for (int q = tid; q < ncr_all * ncc_all; q += K3_THREADS) {
            const uint32_t gb = s_gtc[q], un = s_unc[q];
            c5[2] += __popc(gb);
            c5[3] += __popc(un & gb); c5[4] += __popc(un);
        }
        if (P.uni_mask) {
            // union tile -> bytes.  Thread = (output row, cell): writes the cell's <=4 pixels of that row.
            for (int q = tid; q < nyrows * ncc_all; q += K3_THREADS) {
                const int yr = q / ncc_all, cj = q - yr * ncc_all - 1;
                const int y = y_lo + yr;
                const int ci = (y < 2) ? -1 : (y - 2) >> 2, ry = (y < 2) ? y : (y - 2) & 3;
                const int xbase = (cj < 0) ? 0 : 4 * cj + 2, nrx = (cj < 0) ? 2 : min(4, S_w - xbase);
                const unsigned nib = (s_unc[(ci - ci_lo) * ncc_all + cj + 1] >> (4 * ry)) & 0xfu;
                uint8_t *o = P.uni_mask + ((size_t)b * S_h + y) * S_w + xbase;   // 2-byte aligned
                *reinterpret_cast<uchar2 *>(o) = make_uchar2(nib & 1u, (nib >> 1) & 1u);
                if (nrx == 4) *reinterpret_cast<uchar2 *>(o + 2) = make_uchar2((nib >> 2) & 1u, (nib >> 3) & 1u);
            }
        }
        __syncthreads();   // tiles and tables are released
        BT_PHASE_MARK(2, 5);   // counters + dense outputs
        seq_base = next_base;
        prebuilt = prebuild_next;
        b = nb; s = ns;
    }
    flush_image(cur_b, strips_of_b);
    // When launched as a programmatic dependent (of the COCO matching kernel, whose data this kernel never
    // touches) the completion of this grid must still imply the completion of that one; it finished long ago.
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

static size_t k3_layout(K3Params &P) {
    const int rowsmax = P.R + 1;
    size_t off = (size_t)P.NS * NM * P.PW * sizeof(float);
    P.wpr = P.S_w / 32;
    const size_t celltile = align_up((size_t)rowsmax * (P.PW + 1) * sizeof(uint32_t), 16);
    P.off_lm = (int)off; off += (size_t)rowsmax * P.PW * sizeof(float);
    P.off_scr = (int)off; off += (size_t)(P.scr_cap + rowsmax * P.PW) * sizeof(float);
    P.off_gtrow = (int)off; off += align_up((size_t)(4 * P.R + 2) * (P.wpr + 1) * sizeof(uint32_t), 16);
    P.off_gtc = (int)off; off += celltile;
    P.off_unc = (int)off; off += celltile;
    P.off_list = (int)off; off += 2 * align_up((size_t)P.K * sizeof(unsigned short), 16);
    P.off_cf = (int)off; off += align_up((size_t)ECAP * CF_PITCH * sizeof(float), 16);
    return off;
}

// 3-D tensor map of the prototypes: dims (fastest first) {PW, PH, B*NM}, box {PW, 1, NM}.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_proto_tmap(CUtensorMap *tm, const float *protos, int B, int PH, int PW) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) != cudaSuccess || !sym ||
            qr != cudaDriverEntryPointSuccess)
            return BT_ERR_CUDA;
        fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)PW, (cuuint64_t)PH, (cuuint64_t)B * NM};
    const cuuint64_t gstride[2] = {(cuuint64_t)PW * sizeof(float), (cuuint64_t)PW * PH * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)PW, 1u, (cuuint32_t)NM};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(protos), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? BT_OK : BT_ERR_CUDA;
}

static bool g_pdl = false;   // set by launch_masks for the launch below (host-side, per call)
template <int TPW, int NT, int MINB>
static int launch_k3(const K3Params &P, const CUtensorMap &tm, int grid, size_t smem, cudaStream_t s) {
    if (cudaFuncSetAttribute(masks_kernel<TPW, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return BT_ERR_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pdl ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, masks_kernel<TPW, NT, MINB>, P, tm) != cudaSuccess) return BT_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? BT_OK : BT_ERR_CUDA;
}

int launch_masks(const BtParams &p, const BtIO &io, const Workspace &w, cudaStream_t s, bool pdl) {
    g_pdl = pdl;
    K3Params P{};
    P.B = p.batch; P.S_h = p.img_h; P.S_w = p.img_w; P.PH = p.proto_h; P.PW = p.proto_w;
    P.K = p.max_det; P.crop = p.crop; P.gt_f32 = p.gt_mask_dtype == BT_MASK_F32;
    P.bias = p.proj_bias;
    P.protos = io.protos; P.proj_weight = io.proj_weight; P.det_coeff = io.det_coeff;
    P.det_count = io.det_count; P.masks_gt = io.masks_gt; P.det_region = w.det_region;
    P.strip_done = w.strip_done; P.acc = w.acc; P.inst_area = io.inst_area; P.inst_inter = io.inst_inter;
    P.seg_cnt4 = (long long *)io.seg_cnt4; P.uni_cnt4 = (long long *)io.uni_cnt4;
    P.seg_img3 = (long long *)io.seg_img3; P.uni_img3 = (long long *)io.uni_img3;
    P.seg_dice = io.seg_dice; P.seg_iou = io.seg_iou; P.uni_dice = io.uni_dice; P.uni_iou = io.uni_iou;
    P.seg_mask = io.seg_mask; P.uni_mask = io.uni_mask; P.seg_logits = io.seg_logits;
    P.seg_prob_sum = io.seg_prob_sum;
    if (p.proto_w % 4 != 0 || p.proto_w > 256) return BT_ERR_UNSUPPORTED;   // TMA box width <= 256
    // Two persistent 256-thread CTAs per SM (their phases overlap: the contraction is bound by shared-
    // memory bandwidth, the upsample/threshold by the ALUs) when a ring of R + 1 rows with R >= 2 fits
    // half an SM's shared memory -- a strip releases its rows after its last contraction, so the next
    // strip's rows land during its second half; otherwise one 512-thread CTA with the deepest ring.
    int nt = 0;
    static const bool force_fat = getenv("BTPOST_K3_FAT") != nullptr;   // developer switch (scripts/): compare the two configurations
    for (int R = 3; R >= 2 && !nt && !force_fat; --R) {
        K3Params tmp = P; tmp.R = R; tmp.NS = R + 1; tmp.scr_cap = 1792;
        if ((R + 1) * P.PW <= tmp.scr_cap && k3_layout(tmp) + 128 + 6144 <= 113 * 1024) { P.R = R; P.NS = R + 1; P.scr_cap = 1792; nt = 256; }
    }
    for (int R = 4; R >= 1 && !nt; --R)
        for (int NS = 2 * R + 1; NS >= R + 1 && !nt; --NS) {
            if (NS > NS_MAX) continue;
            K3Params tmp = P; tmp.R = R; tmp.NS = NS; tmp.scr_cap = 4096;
            if (k3_layout(tmp) + 128 <= 222 * 1024) { P.R = R; P.NS = NS; P.scr_cap = 4096; nt = 512; }
        }
    if (!nt) return BT_ERR_UNSUPPORTED;
    const size_t smem = k3_layout(P) + 128;
    P.nstrips = (p.proto_h + P.R - 1) / P.R;
    P.total_strips = P.nstrips * p.batch;
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
            return BT_ERR_CUDA;
    }
    const int ctas = sm_count * (nt == 256 ? 2 : 1);
    const int grid = P.total_strips < ctas ? P.total_strips : ctas;
    CUtensorMap tm;
    if (make_proto_tmap(&tm, io.protos, p.batch, p.proto_h, p.proto_w) != BT_OK) return BT_ERR_CUDA;
    static const char *nt_env = getenv("BTPOST_K3_NT");   // developer switch: threads per CTA of the 2-CTA configuration
    if (nt == 256 && nt_env && P.PW == 160) {
        if (atoi(nt_env) == 384) return launch_k3<160, 384, 2>(P, tm, grid, smem, s);
        if (atoi(nt_env) == 512) return launch_k3<160, 512, 2>(P, tm, grid, smem, s);
    }
    if (nt == 256) {
        if (P.PW == 160) return launch_k3<160, 256, 2>(P, tm, grid, smem, s);
        if (P.PW == 256) return launch_k3<256, 256, 2>(P, tm, grid, smem, s);
        return launch_k3<0, 256, 2>(P, tm, grid, smem, s);
    }
    if (P.PW == 160) return launch_k3<160, 512, 1>(P, tm, grid, smem, s);
    if (P.PW == 256) return launch_k3<256, 512, 1>(P, tm, grid, smem, s);
    return launch_k3<0, 512, 1>(P, tm, grid, smem, s);
}

}  // namespace bt
