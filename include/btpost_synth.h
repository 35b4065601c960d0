/*
 * btpost_synth.h -- device-side generator of the synthetic workload (bench / sweep / test infrastructure, not part
 * of the reference path).  Every element is a pure function of (seed, tensor id, global image index, element index),
 * so any shard materialises its own images without moving data; the tensors are bit-identical to the numpy
 * generator in multitask-bonetumor-yolo_b200/btpost/synth.py (tests/test_gpu_synth.py), which SURVEY.md 8(d)
 * "value distributions & seeds" describes.
 */
#ifndef BTPOST_SYNTH_H_
#define BTPOST_SYNTH_H_

#include <stdint.h>

#include "btpost.h"

#ifdef __cplusplus
extern "C" {
#endif

/* keys_head / keys_proto: [B] u64 stream keys of the images (host-computed: synth.stream_key);
 * objects: [B, 3, 6] fp32 rows (valid, cls, cx, cy, w, h) in pixels (synth.object_table);
 * outputs (device): head [B, 4+nc+nm, N] fp32, protos [B, nm, S/4, S/4] fp32, masks_gt [B, 1, S, S] u8.
 * Any output pointer may be NULL.  All pointers are device pointers; work is enqueued on `stream`. */
BTPOST_API int btpost_synth_batch(int32_t batch, int32_t img_size, int32_t nc, int32_t nm, const uint64_t *keys_head,
                                  const uint64_t *keys_proto, const float *objects, float *head, float *protos,
                                  uint8_t *masks_gt, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BTPOST_SYNTH_H_ */
