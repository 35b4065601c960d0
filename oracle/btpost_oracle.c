/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar, one thread) of the reference's
 * post-processing + evaluation hot path.  It is the checker for the CUDA library: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  Nothing in the product
 * (multitask-bonetumor-yolo_b200/) links, imports or calls it.
 *
 * Every function cites the reference statement it follows (paths under /root/reference/).
 * Third-party arithmetic that is not vendored there is restated from its published algorithm
 * and pinned in tests/ against the installed torchvision 0.26.0 CPU `nms` and against fixtures
 * produced by running the reference's own validation_step (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -mfma -ffp-contract=off (see oracle/Makefile).  With
 * -ffp-contract=off the compiler never fuses a*b+c on its own; every fused multiply-add below is
 * an explicit fmaf() and every other operation rounds once to fp32 -- the CUDA kernels use the
 * same sequence (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn), so results are bit-identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BTO_API __attribute__((visibility("default")))

/* sigmoid(x) > 0.5 evaluated in fp32 as torch does (src/running_main_v2.py:702-703,
 * src/test_model.py:85) is true exactly for x >= 0x33c00001: pinned by bisection over fp32 bit
 * patterns against torch 2.11 CPU (tests/golden/make_golden.py::pin_sigmoid_threshold). */
static const float BTO_SIGMOID_HALF_X = 8.940696716308594e-08f; /* 1.5 * 2^-24 = 0x33c00000 */

/* ------------------------------------------------------------------ decode (a2, L2 form) */
/* L2 `segment_preds_cat` [4+nc+nm, N] (src/main_modelv2.py:367-375): rows 0..3 are xywh in
 * pixels; decode = Ultralytics xywh2xyxy (x1 = cx - w/2, x2 = cx + w/2);
 * score/label = max / first argmax over class rows (src/running_main_v2.py:788). */
BTO_API void bto_decode_l2(const float *head, int N, int nc, float *boxes, float *score, int32_t *label) {
    const float *cx = head, *cy = head + N, *w = head + 2 * (size_t)N, *h = head + 3 * (size_t)N;
    for (int n = 0; n < N; ++n) {
        float hw = w[n] * 0.5f, hh = h[n] * 0.5f;
        boxes[4 * n + 0] = cx[n] - hw;
        boxes[4 * n + 1] = cy[n] - hh;
        boxes[4 * n + 2] = cx[n] + hw;
        boxes[4 * n + 3] = cy[n] + hh;
        float best = head[(size_t)4 * N + n];
        int32_t bi = 0;
        for (int c = 1; c < nc; ++c) {
            float s = head[(size_t)(4 + c) * N + n];
            /* torch .max(dim): a NaN is the maximum and the first NaN wins (pinned in tests/test_oracle_pinning.py) */
            if (s > best || (s != s && best == best)) { best = s; bi = c; }
        }
        score[n] = best;
        label[n] = bi;
    }
}

/* ------------------------------------------------------------------ decode (a2, L1 form) */
/* Shared exp: Cody-Waite range reduction + degree-6 polynomial, only fmaf/mul/add, so the C and
 * CUDA sides agree bit-for-bit (|rel err| < 2 ulp vs libm on [-88, 0]). */
static inline float bto_expf(float x) {
    if (x < -87.0f) return 0.0f;
    float t = x * 1.4426950408889634f;
    float n = nearbyintf(t);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.428606765330187e-06f, r);
    float p = 1.3888889225e-03f;
    p = fmaf(p, r, 8.3333337680e-03f);
    p = fmaf(p, r, 4.1666667908e-02f);
    p = fmaf(p, r, 1.6666667163e-01f);
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int32_t e = (int32_t)n;
    union { float f; int32_t i; } u;
    u.f = p;
    u.i += e << 23;
    return u.f;
}

BTO_API float bto_expf_public(float x) { return bto_expf(x); }

/* One level of the raw maps [4*R+nc, H, W] (src/running_main_v2.py:743-775; dist2bbox :97-107):
 * softmax over R bins -> expectation with arange(R) -> (anchor -/+ ltrb) * stride, anchors
 * (x+0.5, y+0.5) row-major, stride = img/W; class score = sigmoid(logit).  Output in xyxy. */
BTO_API void bto_decode_l1_level(const float *map, int H, int W, int R, int nc, float stride,
                                 float *boxes, float *scores /* [HW, nc] */) {
    int HW = H * W;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int n = y * W + x;
            float d[4];
            for (int s = 0; s < 4; ++s) {
                const float *p = map + (size_t)(s * R) * HW + n;
                float m = p[0];
                for (int k = 1; k < R; ++k) { float v = p[(size_t)k * HW]; if (v > m) m = v; }
                float sum = 0.0f, acc = 0.0f;
                for (int k = 0; k < R; ++k) {
                    float e = bto_expf(p[(size_t)k * HW] - m);
                    sum += e;
                    acc = fmaf(e, (float)k, acc);
                }
                d[s] = acc / sum;
            }
            float ax = ((float)x + 0.5f) * stride, ay = ((float)y + 0.5f) * stride;
            boxes[4 * n + 0] = ax - d[0] * stride;
            boxes[4 * n + 1] = ay - d[1] * stride;
            boxes[4 * n + 2] = ax + d[2] * stride;
            boxes[4 * n + 3] = ay + d[3] * stride;
            for (int c = 0; c < nc; ++c) {
                float l = map[(size_t)(4 * R + c) * HW + n];
                scores[(size_t)n * nc + c] = 1.0f / (1.0f + bto_expf(-l));
            }
        }
}

/* ------------------------------------------------------------------ filter (a3) */
/* src/running_main_v2.py:788-795 / running_main_v3.py:538-548: keep = top_score > CONF_TH
 * (strict), boolean gather in anchor order, clamp_(0, img_size) after the gather. */
BTO_API int bto_filter(const float *boxes, const float *score, const int32_t *label, int N, float conf,
                       int clamp, float img_w, float img_h, float *cbox, float *cscore,
                       int32_t *clabel, int32_t *canchor) {
    int m = 0;
    for (int n = 0; n < N; ++n) {
        if (!(score[n] > conf)) continue;
        for (int k = 0; k < 4; ++k) {
            float v = boxes[4 * n + k];
            if (clamp) {
                float hi = (k & 1) ? img_h : img_w;
                v = v < 0.0f ? 0.0f : (v > hi ? hi : v);
            }
            cbox[4 * m + k] = v;
        }
        cscore[m] = score[n];
        clabel[m] = label[n];
        canchor[m] = n;
        ++m;
    }
    return m;
}

/* ------------------------------------------------------------------ NMS (a4) */
/* torchvision.ops.nms CPU (un-vendored third party; call sites src/running_main_v2.py:817,
 * src/running_main_v3.py:549; algorithm restated from SURVEY.md A.1 and pinned against the
 * installed torchvision 0.26.0 in tests/test_oracle_pinning.py): stable descending sort (NaN
 * first, ties -> lower index), greedy, fp32 IoU inter/(a_i + a_j - inter), suppress iff
 * (double)iou > iou_threshold, std::max/min operand order kept so NaN behaves the same. */
typedef struct { uint32_t key; int32_t idx; } bto_sortrec;

static inline uint32_t bto_desc_key(float s) {
    union { float f; uint32_t u; } v;
    v.f = s;
    if (s != s) return 0u;                                         /* NaN sorts first */
    uint32_t asc = (v.u & 0x80000000u) ? ~v.u : (v.u | 0x80000000u); /* ascending-orderable */
    if (v.u == 0x80000000u) asc = 0x80000000u;                     /* -0.0 == +0.0 */
    return ~asc;                                                   /* smaller = higher score */
}

static int bto_cmp(const void *a, const void *b) {
    const bto_sortrec *x = (const bto_sortrec *)a, *y = (const bto_sortrec *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static inline float bto_max(float a, float b) { return (a < b) ? b : a; } /* std::max(a,b) */
static inline float bto_min(float a, float b) { return (b < a) ? b : a; } /* std::min(a,b) */

/* class_mode 0: agnostic (reference).  1: class-aware, same-class predicate on raw coordinates
 * (== per-class NMS merged by score).  2: Ultralytics offset boxes + label*max_wh
 * in fp32, then agnostic. */
BTO_API int bto_nms(const float *boxes_in, const float *scores, const int32_t *labels, int n,
                    double iou_thr, int class_mode, float max_wh, int64_t *keep, int max_keep, int max_cand) {
    if (n <= 0 || max_keep <= 0) return 0;
    float *boxes = (float *)malloc(sizeof(float) * 4 * (size_t)n);
    float *area = (float *)malloc(sizeof(float) * (size_t)n);
    bto_sortrec *ord = (bto_sortrec *)malloc(sizeof(bto_sortrec) * (size_t)n);
    uint8_t *sup = (uint8_t *)calloc((size_t)n, 1);
    for (int i = 0; i < n; ++i) {
        float off = (class_mode == 2 && labels) ? (float)labels[i] * max_wh : 0.0f;
        for (int k = 0; k < 4; ++k) boxes[4 * i + k] = (class_mode == 2) ? boxes_in[4 * i + k] + off : boxes_in[4 * i + k];
        area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
        ord[i].key = bto_desc_key(scores[i]);
        ord[i].idx = i;
    }
    qsort(ord, (size_t)n, sizeof(bto_sortrec), bto_cmp);
    /* Ultralytics non_max_suppression: x = x[x[:, 4].argsort(descending=True)[:max_nms]] before torchvision.ops.nms
     * (SURVEY.md A.5).  argsort is made stable here (ties -> lower index), as the NMS sort itself is. */
    if (max_cand > 0 && max_cand < n) n = max_cand;
    int nk = 0;
    for (int _i = 0; _i < n && nk < max_keep; ++_i) {
        int i = ord[_i].idx;
        if (sup[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float ia = area[i];
        for (int _j = _i + 1; _j < n; ++_j) {
            int j = ord[_j].idx;
            if (sup[j]) continue;
            if (class_mode == 1 && labels && labels[i] != labels[j]) continue;
            float xx1 = bto_max(ix1, boxes[4 * j]), yy1 = bto_max(iy1, boxes[4 * j + 1]);
            float xx2 = bto_min(ix2, boxes[4 * j + 2]), yy2 = bto_min(iy2, boxes[4 * j + 3]);
            float w = bto_max(0.0f, xx2 - xx1), h = bto_max(0.0f, yy2 - yy1);
            float inter = w * h;
            float ovr = inter / (ia + area[j] - inter);
            if ((double)ovr > iou_thr) sup[j] = 1;
        }
    }
    free(boxes); free(area); free(ord); free(sup);
    return nk;
}

/* ------------------------------------------------------------------ GT prep (a6) */
/* src/running_main_v2.py:842-882 (mAP copy, clamped) and :403-433 (loss copy, unclamped).
 * mode 0 reproduces the shipped statement literally: torch.cat([x1[G], y1[G], x2[G], y2[G]],
 * dim=-1).view(-1, 4) -- a flat [4G] vector re-read row-major, so for G >= 2 box g, column c
 * holds coord[(4g+c)/G][(4g+c)%G] (a reference defect).  mode 1 is the
 * per-object xyxy the author intended. */
BTO_API int bto_gt_prep(const float *gt, int G_total, int b, float S, int mode, int clamp,
                        float *out_boxes, int32_t *out_labels, int max_gt) {
    int g = 0;
    float (*coord)[4] = malloc(sizeof(float[4]) * (size_t)(G_total > 0 ? G_total : 1));
    for (int r = 0; r < G_total; ++r) {
        const float *row = gt + 6 * (size_t)r;
        if (!(row[0] == (float)b)) continue;
        if (g >= max_gt) break;
        float cx = row[2], cy = row[3], w = row[4], h = row[5];
        coord[g][0] = (cx - w / 2.0f) * S;
        coord[g][1] = (cy - h / 2.0f) * S;
        coord[g][2] = (cx + w / 2.0f) * S;
        coord[g][3] = (cy + h / 2.0f) * S;
        out_labels[g] = (int32_t)row[1];
        ++g;
    }
    for (int i = 0; i < 4 * g; ++i) {
        float v = (mode == 0) ? coord[i % g][i / g] : coord[i / 4][i % 4];
        if (clamp) v = v < 0.0f ? 0.0f : (v > S ? S : v);
        out_boxes[i] = v;
    }
    free(coord);
    return g;
}

/* ------------------------------------------------------------------ CM matching (a10) */
/* src/running_main_v2.py:435-449,476-486 + batch_bbox_iou :68-94: IoU of all N raw decoded
 * (unclamped) boxes vs the image's GT (loss copy), iou = inter/(union + 1e-7) in fp32, max and
 * first argmax over GT, positive iff > iou_match_thresh; pair = (argmax class, gt class).
 * cm is [nc, nc] int64 indexed [gt_class][pred_class] (MulticlassConfusionMatrix [true, pred],
 * SURVEY.md A.4). Returns number of positive anchors. */
BTO_API int bto_cm_match(const float *boxes, const int32_t *pred_cls, int N, const float *gtb,
                         const int32_t *gtl, int G, float thr, int nc, int64_t *cm, int32_t *pos_anchor) {
    int npos = 0;
    if (G <= 0) return 0;
    for (int n = 0; n < N; ++n) {
        const float *p = boxes + 4 * (size_t)n;
        float a1 = (p[2] - p[0]) * (p[3] - p[1]);
        float best = -INFINITY; int bi = 0;
        for (int g = 0; g < G; ++g) {
            const float *q = gtb + 4 * g;
            float x1 = fmaxf(p[0], q[0]), y1 = fmaxf(p[1], q[1]);
            float x2 = fminf(p[2], q[2]), y2 = fminf(p[3], q[3]);
            float iw = x2 - x1; iw = iw < 0.0f ? 0.0f : iw;
            float ih = y2 - y1; ih = ih < 0.0f ? 0.0f : ih;
            float inter = iw * ih;
            float a2 = (q[2] - q[0]) * (q[3] - q[1]);
            float uni = a1 + a2 - inter;
            float iou = inter / (uni + 1e-7f);
            if (g == 0 || iou > best) { best = iou; bi = g; }
        }
        if (best > thr) {
            int gc = gtl[bi], pc = pred_cls[n];
            if (gc >= 0 && gc < nc && pc >= 0 && pc < nc) cm[(size_t)gc * nc + pc] += 1;
            if (pos_anchor) pos_anchor[npos] = n;
            ++npos;
        }
    }
    return npos;
}

/* ------------------------------------------------------------------ mask assembly (a7) */
/* M1 projector: Conv2d(nm->1, k=1) (src/running_main_v2.py:197,689-691) = bias + sum_k w_k p_k;
 * sequential fmaf from the bias reproduces torch CPU conv2d bit-for-bit (pinned in
 * make_golden.py).  M2 instance: coeff . protos (src/test_model.py:81; Ultralytics
 * process_mask `masks_in @ protos`), sequential fmaf from 0 (== MKL sgemm at K=32, pinned). */
BTO_API void bto_project(const float *protos, const float *w, float bias, int nm, int hw, float *logit) {
    for (int i = 0; i < hw; ++i) {
        float acc = bias;
        for (int k = 0; k < nm; ++k) acc = fmaf(w[k], protos[(size_t)k * hw + i], acc);
        logit[i] = acc;
    }
}

/* F.interpolate(mode="bilinear", align_corners=False) with an explicit size
 * (src/running_main_v2.py:694-699).  torch CPU evaluates
 * fma(h0, fma(w0, v00, w1*v01), h1 * fma(w0, v10, w1*v11)) -- pinned bit-exact in make_golden.py. */
static void bto_axis(int in, int out, int d, int *i0, int *i1, float *l0, float *l1) {
    float scale = (float)in / (float)out;
    float src = scale * ((float)d + 0.5f) - 0.5f;
    if (src < 0.0f) src = 0.0f;
    int f = (int)src;
    *i0 = f;
    *i1 = f + 1 < in ? f + 1 : in - 1;
    *l1 = src - (float)f;
    *l0 = 1.0f - *l1;
}

static inline float bto_bilerp(float v00, float v01, float v10, float v11, float w0, float w1, float h0, float h1) {
    float r0 = fmaf(w0, v00, w1 * v01);
    float r1 = fmaf(w0, v10, w1 * v11);
    return fmaf(h0, r0, h1 * r1);
}

BTO_API void bto_upsample(const float *in, int ih, int iw, float *out, int oh, int ow) {
    for (int y = 0; y < oh; ++y) {
        int y0, y1; float h0, h1;
        bto_axis(ih, oh, y, &y0, &y1, &h0, &h1);
        for (int x = 0; x < ow; ++x) {
            int x0, x1; float w0, w1;
            bto_axis(iw, ow, x, &x0, &x1, &w0, &w1);
            out[(size_t)y * ow + x] = bto_bilerp(in[y0 * iw + x0], in[y0 * iw + x1], in[y1 * iw + x0],
                                                 in[y1 * iw + x1], w0, w1, h0, h1);
        }
    }
}

BTO_API void bto_threshold(const float *logit, size_t n, uint8_t *mask) {
    for (size_t i = 0; i < n; ++i) mask[i] = logit[i] > BTO_SIGMOID_HALF_X;
}

/* Instance mask (north-star M2): contraction -> crop at prototype resolution (Ultralytics
 * crop_mask: keep r >= x1 & r < x2 & c >= y1 & c < y2 on boxes scaled by pw/iw, ph/ih;
 * SURVEY.md A.5) -> bilinear upsample of the logits -> sigmoid > 0.5 (src/test_model.py:82-85). */
BTO_API void bto_instance_mask(const float *protos, const float *coeff, const float *box, int nm,
                               int ph, int pw, int ih, int iw, int crop, float *scratch, uint8_t *mask) {
    int hw = ph * pw;
    float rx = (float)((double)pw / (double)iw), ry = (float)((double)ph / (double)ih);
    float x1 = box[0] * rx, y1 = box[1] * ry, x2 = box[2] * rx, y2 = box[3] * ry;
    for (int r = 0; r < ph; ++r)
        for (int c = 0; c < pw; ++c) {
            int i = r * pw + c;
            float acc = 0.0f;
            for (int k = 0; k < nm; ++k) acc = fmaf(coeff[k], protos[(size_t)k * hw + i], acc);
            if (crop) {
                int in = ((float)c >= x1) && ((float)c < x2) && ((float)r >= y1) && ((float)r < y2);
                acc = acc * (in ? 1.0f : 0.0f);
            }
            scratch[i] = acc;
        }
    for (int y = 0; y < ih; ++y) {
        int y0, y1i; float h0, h1;
        bto_axis(ph, ih, y, &y0, &y1i, &h0, &h1);
        for (int x = 0; x < iw; ++x) {
            int x0, x1i; float w0, w1;
            bto_axis(pw, iw, x, &x0, &x1i, &w0, &w1);
            float v = bto_bilerp(scratch[y0 * pw + x0], scratch[y0 * pw + x1i], scratch[y1i * pw + x0],
                                 scratch[y1i * pw + x1i], w0, w1, h0, h1);
            mask[(size_t)y * iw + x] = v > BTO_SIGMOID_HALF_X;
        }
    }
}

/* ------------------------------------------------------------------ seg counters (a8) */
/* Binarised prediction vs GT (src/running_main_v2.py:702-713): global tp/fp/fn/tn (BinaryF1 /
 * Precision / Recall / Accuracy states) and per-image inter, |P|, |G|
 * (DiceScore; src/test_model.py:15-23).  cnt4 = {tp, fp, fn, tn} is accumulated. */
BTO_API void bto_mask_counts(const uint8_t *pred, const uint8_t *gt, size_t n, int64_t *cnt4, int64_t *img3) {
    int64_t tp = 0, p = 0, g = 0;
    for (size_t i = 0; i < n; ++i) {
        int a = pred[i] != 0, b = gt[i] != 0;
        tp += a & b; p += a; g += b;
    }
    cnt4[0] += tp; cnt4[1] += p - tp; cnt4[2] += g - tp; cnt4[3] += (int64_t)n - p - g + tp;
    img3[0] = tp; img3[1] = p; img3[2] = g;
}

/* Per-image Dice / IoU exactly as src/test_model.py:15-23 evaluates them in fp32:
 * iou = (inter + eps) / (union + eps); dice = (2*inter + eps) / (|P| + |G| + eps). */
BTO_API void bto_dice_iou(int64_t inter, int64_t p, int64_t g, float *dice, float *iou) {
    float fi = (float)inter, fu = (float)(p + g - inter);
    *iou = (fi + 1e-7f) / (fu + 1e-7f);
    *dice = (2.0f * fi + 1e-7f) / ((float)p + (float)g + 1e-7f);
}

/* ------------------------------------------------------------------ COCO matching (a9) */
/* torchmetrics MeanAveragePrecision -> pycocotools COCOeval.evaluateImg for one (image, class),
 * restated from SURVEY.md A.3 (pycocotools is not installed: parity unpinned).  Boxes arrive as
 * fp32 xyxy; pycocotools sees xywh with w = x2 - x1, h = y2 - y1 computed in fp32 and widened.
 * dets must be in descending-score order (they are: NMS output order).  For each area range a
 * and IoU threshold t writes dt_match[a][t][d] (matched gt index + 1, 0 = none) and
 * dt_ignore[a][t][d]; gt_ignore[a][g]. */
static double bto_bbiou(const double *d, const double *g) {
    double w = fmin(d[0] + d[2], g[0] + g[2]) - fmax(d[0], g[0]);
    if (w <= 0) return 0.0;
    double h = fmin(d[1] + d[3], g[1] + g[3]) - fmax(d[1], g[1]);
    if (h <= 0) return 0.0;
    double i = w * h;
    return i / (d[2] * d[3] + g[2] * g[3] - i);
}

BTO_API void bto_coco_match(const float *det_xyxy, int D, const float *gt_xyxy, int G,
                            const double *iou_thrs, int T, const double *area_rng /* [A][2] */, int A,
                            int32_t *dt_match, uint8_t *dt_ignore, uint8_t *gt_ignore) {
    double *db = malloc(sizeof(double) * 4 * (size_t)(D > 0 ? D : 1));
    double *gb = malloc(sizeof(double) * 4 * (size_t)(G > 0 ? G : 1));
    double *ious = malloc(sizeof(double) * (size_t)(D > 0 ? D : 1) * (size_t)(G > 0 ? G : 1));
    int *gord = malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int *gm = malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    for (int d = 0; d < D; ++d) {
        const float *b = det_xyxy + 4 * d;
        db[4 * d] = b[0]; db[4 * d + 1] = b[1]; db[4 * d + 2] = (double)(b[2] - b[0]); db[4 * d + 3] = (double)(b[3] - b[1]);
    }
    for (int g = 0; g < G; ++g) {
        const float *b = gt_xyxy + 4 * g;
        gb[4 * g] = b[0]; gb[4 * g + 1] = b[1]; gb[4 * g + 2] = (double)(b[2] - b[0]); gb[4 * g + 3] = (double)(b[3] - b[1]);
    }
    for (int d = 0; d < D; ++d)
        for (int g = 0; g < G; ++g) ious[(size_t)d * G + g] = bto_bbiou(db + 4 * d, gb + 4 * g);
    for (int a = 0; a < A; ++a) {
        double lo = area_rng[2 * a], hi = area_rng[2 * a + 1];
        uint8_t *gi = gt_ignore + (size_t)a * G;
        for (int g = 0; g < G; ++g) { double ar = gb[4 * g + 2] * gb[4 * g + 3]; gi[g] = (ar < lo || ar > hi); }
        int k = 0;                                    /* stable: non-ignored first */
        for (int g = 0; g < G; ++g) if (!gi[g]) gord[k++] = g;
        for (int g = 0; g < G; ++g) if (gi[g]) gord[k++] = g;
        for (int t = 0; t < T; ++t) {
            int32_t *dm = dt_match + ((size_t)a * T + t) * D;
            uint8_t *di = dt_ignore + ((size_t)a * T + t) * D;
            for (int g = 0; g < G; ++g) gm[g] = 0;
            for (int d = 0; d < D; ++d) {
                double best = iou_thrs[t] < 1 - 1e-10 ? iou_thrs[t] : 1 - 1e-10;
                int m = -1;
                for (int s = 0; s < G; ++s) {
                    int g = gord[s];
                    if (gm[g]) continue;
                    if (m > -1 && !gi[gord[m]] && gi[g]) break;
                    double v = ious[(size_t)d * G + g];
                    if (v < best) continue;
                    best = v; m = s;
                }
                if (m >= 0) {
                    int g = gord[m];
                    gm[g] = 1; dm[d] = g + 1; di[d] = gi[g];
                } else {
                    double ar = db[4 * d + 2] * db[4 * d + 3];
                    dm[d] = 0; di[d] = (ar < lo || ar > hi);
                }
            }
        }
    }
    free(db); free(gb); free(ious); free(gord); free(gm);
}
