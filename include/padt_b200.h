/* padt_b200 -- C ABI of the B200-native PaDT generate/decode hot path.
 *
 * The reference (Gorilla-Lab-SCUT/PaDT) has no FFI of its own: its boundary is three Python names
 * (src/PaDT/__init__.py:1).  This library is what those names call into in padt_b200/ (ctypes binding in
 * padt_b200/_capi.py); each entry point below cites the reference code it replaces.  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions: all data pointers are BORROWED device pointers (torch-owned storage) unless a name ends in _host;
 * activations/weights are bf16, positional tables fp32, ids int64; every call takes a cudaStream_t, returns 0 on
 * success or <0 (message via padt_last_error()); one handle per (process, GPU); a handle is not thread-safe.
 */
#ifndef PADT_B200_H
#define PADT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PadtHandle PadtHandle;
typedef void* padt_stream_t; /* cudaStream_t */

/* Model dimensions == padt_b200.config.PadtConfig (Qwen2.5-VL config.json + the `vl_decoder` dict of
 * src/PaDT/trainer/padt_sft_trainer.py:151-160 as fixed up at src/PaDT/models/padt.py:127-130). */
typedef struct PadtDims {
    int32_t vit_depth, vit_hidden, vit_inter, vit_heads, vit_in_ch, vit_patch, vit_tpatch, vit_merge, vit_window, vit_out_hidden;
    int32_t vit_n_fullatt;
    int32_t vit_fullatt[8];
    int32_t D, I, n_layers, n_heads, n_kv_heads, vocab;
    float rms_eps;
    int32_t mrope0, mrope1, mrope2;
    int32_t tie_word_embeddings;
    int32_t dec_hidden, dec_inter, dec_heads;
    int32_t use_proto_proj, lora_r;
} PadtDims;

const char* padt_last_error(void);
const char* padt_version(void);
/* Number of CUDA kernels this library has launched in this process (graph replays count their kernel nodes). */
unsigned long long padt_launch_count(void);

/* Canonical weight order (== padt_b200.weights.weight_shapes(cfg), HF-4.50 key layout of the reference checkpoints). */
int padt_weight_count(const PadtDims* dims);
const char* padt_weight_name(const PadtDims* dims, int index);

/* Create an engine.  `weights[i]` = device pointer of tensor padt_weight_name(dims, i) (bf16, contiguous).
 * `llm_cos`/`llm_sin`: bf16 [max_pos, head_dim/2] rotary tables (HF:595-608 evaluated on the host).
 * Replaces PaDTForConditionalGeneration.__init__/from_pretrained module construction (padt.py:115-132): the engine
 * packs fused/interleaved copies of the projection weights it streams (q|k|v rows, gate/up interleaved). */
int padt_create(const PadtDims* dims, const void* const* weights, int n_weights, const void* llm_cos, const void* llm_sin,
                int max_pos, PadtHandle** out);
int padt_destroy(PadtHandle* h);
/* Replace the LLM rotary tables (same layout as in padt_create) when a request needs more positions than the current table
 * holds: the reference evaluates cos/sin per call and has no length limit (HF:595-608).  The caller keeps the tensors alive. */
int padt_set_rope_table(PadtHandle* h, const void* llm_cos, const void* llm_sin, int max_pos);

/* V1-V6: custom_visual_forward (padt.py:48-106).  pixel_values [P,1176]; grid_thw_host [n_img*3];
 * pos_cos/pos_sin: fp32 DEVICE tables [tab_rows, head_dim/4] = cos/sin(outer(arange, inv_freq)) (HF:117-130).
 * Outputs: image_embeds [P/4, D] (raster-merged order), high_res [P, vit_hidden] (window order), cos/sin fp32
 * [P, head_dim] (window order). */
int padt_vit_forward(PadtHandle* h, const void* pixel_values, const int64_t* grid_thw_host, int n_img, const float* pos_cos,
                     const float* pos_sin, int tab_rows, void* image_embeds, void* high_res, float* cos, float* sin,
                     padt_stream_t stream);

/* V3/V4 host integer tables only (no GPU needed): rot_pos_emb / get_window_index (HF:382-453) as the engine builds them.
 * embed_row_map[P]: window-order row of each input patch; hpos/wpos[P]: rotary ids in window order; merge_row_map[M] ==
 * window_index; win_len[n_win] patches per (non-empty) window.  Any output pointer may be NULL. */
int padt_vit_plan_host(const PadtDims* dims, const int64_t* grid_thw_host, int n_img, int32_t* embed_row_map, int32_t* hpos,
                       int32_t* wpos, int32_t* merge_row_map, int32_t* win_len, int32_t* n_win, int32_t* n_patches);

/* E1: prototype projection (padt.py:187-191). */
int padt_prototypes(PadtHandle* h, const void* image_embeds, int n_rows, void* prototypes, padt_stream_t stream);

/* Generation-time buffers the caller owns (all device). */
typedef struct PadtGenBuffers {
    int64_t* tokens_out;    /* [max_steps,B] tokens appended to the sequences (padt.py:749-752) */
    int64_t* argmax_out;    /* [max_steps,B] raw arg-max per step, may be NULL */
    void* hidden_steps;     /* [max_steps,B,D] final-norm state that predicted token k (padt.py:732-737), may be NULL */
    void* last_hidden;      /* [B,D] */
    void* prefill_hidden;   /* [B,L,D] final-norm states of the prompt (hidden_states[0][-1]), may be NULL */
    const int64_t* forced;  /* [max_steps,B] teacher-forced feedback tokens, may be NULL.  The sampler reads entry t when it samples token t,
                             * so a caller that picks tokens itself (do_sample / logits processors, padt.py:717,740-745) writes entry t on the
                             * stream before the decode call that samples it. */
    int32_t max_steps;
    void* logits_out;       /* [B, logits_ld] bf16: the masked, bf16-rounded extended logits of the latest forward (padt.py:292-301,713), may be
                             * NULL.  Columns no row may emit are not computed: the caller pre-fills the buffer with -inf. */
    int32_t logits_ld;
} PadtGenBuffers;

/* E2-E4, R1, T1, T2 (prefill): forward_main pixel_values branch after the ViT (padt.py:193-219,256-301).
 * input_ids [B,L] int64 (global VRT ids); image_slot [B,L] int32 (k-th <|image_pad|> -> k, else -1: the masked_scatter
 * order of padt.py:219); pos3 [3,B,L] int32 = get_rope_index (transformers 4.50 semantics); *_host arrays have B
 * entries: kv_start = number of left pads, rope_delta (padt.py:266), vrt_lo/hi = prototype-row range the row may emit
 * (padt.py:196-201).  Leaves the arg-max partials of the last position in the handle; call padt_decode next. */
int padt_prefill(PadtHandle* h, const int64_t* input_ids, const int32_t* image_slot, const int32_t* pos3,
                 const int32_t* kv_start_host, const int32_t* rope_delta_host, const int32_t* vrt_lo_host, const int32_t* vrt_hi_host,
                 const void* image_embeds, const void* prototypes, int n_protos, int B, int L,
                 const int64_t* eos_ids_host, int n_eos, int64_t pad_id, const PadtGenBuffers* bufs, padt_stream_t stream);

/* S1 + T1/T2 (decode): n_steps iterations of {sample token t, embed it, run the layers, logits} with no host sync
 * (padt.py:671-757 minus the Python loop), then one trailing sample if `finish` != 0. */
int padt_decode(PadtHandle* h, int n_steps, int finish, padt_stream_t stream);
/* Synchronises the stream and reads {tokens sampled so far, rows still unfinished}. */
int padt_decode_status(PadtHandle* h, int32_t* n_sampled, int32_t* n_unfinished, padt_stream_t stream);
/* Run-time options: "decode_impl" 1 = persistent one-launch-per-token kernel (decode_mega.cu, default; falls back to 0 when the
 * model does not fit its shared-memory plan), 0 = per-layer kernels (decode.cu); "use_graph" 1 = replay the step from a CUDA graph
 * (default), 0 = eager launches; "gemm_tc" 1 = persistent tcgen05 GEMM (default), 2 = one-tile tcgen05 GEMM, 0 = mma.sync GEMM. */
int padt_set_option(PadtHandle* h, const char* name, int value);

/* D1-D4: vl_decode + PaDTDecoder.forward (padt.py:342-412, padt_decoder.py:187-276).
 * vrt_feats [sum n_vrt, D]; obj_vrt_count_host / obj_sample_host [n_obj]; low [sum M, D]; high [sum P, vit_hidden];
 * cos/sin fp32 [sum P, head_dim]; grid_thw_host [n_img*3].  Outputs: boxes [n_obj,4] (sigmoid applied), scores [n_obj,1]
 * (logits), masks [n_obj, 4*hmax, 4*wmax] (logits, zero outside each object's valid region). */
int padt_vl_decode(PadtHandle* h, const void* vrt_feats, const int32_t* obj_vrt_count_host, const int32_t* obj_sample_host, int n_obj,
                   const void* low, const void* high, const float* cos, const float* sin, const int64_t* grid_thw_host, int n_img,
                   void* boxes, void* scores, void* masks, int hmax, int wmax, padt_stream_t stream);

/* Kernel-level entry points used by the parity tests (tests/test_ops_gpu.py). */
int padt_op_gemm(int epilogue, const void* A, int lda, const void* W, int ldw, const void* bias, const void* resid, int ldr, void* C,
                 int ldc, const int32_t* out_row_map, int M, int N, int K, padt_stream_t stream);
int padt_op_rmsnorm(const void* x, const void* w, void* y, int rows, int dim, float eps, padt_stream_t stream);
int padt_op_attention(const void* q, const void* k, const void* v, void* o, int64_t q_row, int64_t q_head, int64_t k_row, int64_t k_head,
                      int64_t v_row, int64_t v_head, int64_t o_row, int64_t o_head, const int32_t* q_off, const int32_t* q_len,
                      const int32_t* k_off, const int32_t* k_len, const int32_t* k_start, int n_seg, int n_heads, int group,
                      int max_q_len, int head_dim, int causal, float scale, padt_stream_t stream);
/* padt.py:301: logits.masked_fill_(~logit_mask[:, None, :], -inf) for the teacher-forced forward: logits [B*rows_per_sample, ld] bf16, columns
 * >= vocab outside [vocab + vrt_lo[b], vocab + vrt_hi[b]) of sample b become -inf.  B <= 16. */
int padt_op_logit_mask(void* logits, int64_t ld, int rows_per_sample, int B, int vocab, int n_ext, const int32_t* vrt_lo_host,
                       const int32_t* vrt_hi_host, padt_stream_t stream);
/* Same as padt_op_attention with the key range of every (segment, head) cut into n_split parts (split-KV, then a merge kernel): the
 * small-q / large-kv cross attention of the PaDT decoder (padt_decoder.py:55 with 3 + n_vrt queries against 9216 image tokens).
 * max_q_len <= 64, non-causal.  Workspaces (device, fp32): part_o [n_seg*n_heads*n_split*64*head_dim], part_ml [n_seg*n_heads*n_split*64*2]. */
int padt_op_attention_split(const void* q, const void* k, const void* v, void* o, int64_t q_row, int64_t q_head, int64_t k_row, int64_t k_head,
                            int64_t v_row, int64_t v_head, int64_t o_row, int64_t o_head, const int32_t* q_off, const int32_t* q_len,
                            const int32_t* k_off, const int32_t* k_len, int n_seg, int n_heads, int max_q_len, int head_dim, float scale,
                            int n_split, float* part_o, float* part_ml, padt_stream_t stream);
int padt_op_rope_neox(void* x, int64_t row_stride, int n_heads, int head_dim, const float* cos, const float* sin, int cs_ld, int rows,
                      padt_stream_t stream);

/* ---- image pre-processing on the GPU.  Replaces the host-side image-processor call of every request
 * (eval/test_demo.py:87-93, eval/evaluation_scripts/utils.py:221-229 -> transformers 4.50 Qwen2-VL image processor: smart_resize, PIL bicubic
 * resize on uint8, * 1/255, (x - mean) / std, patchify).  Bit-exact with Pillow's integer resampler and the float32 arithmetic of the
 * processor; no model handle needed. */
/* smart_resize (image_processing_qwen2_vl.py:62-88): out_hw = {resized_height, resized_width}; error when the aspect ratio exceeds 200. */
int padt_smart_resize(int height, int width, int factor, int64_t min_pixels, int64_t max_pixels, int32_t* out_hw);
size_t padt_preprocess_workspace_bytes(int H, int W, int out_h, int out_w);
/* rgb: device uint8 [H, W, 3] -> pixel_values: device [(out_h / patch) * (out_w / patch), 3 * temporal * patch * patch], bf16 (out_f32 = 0:
 * what the patch-embed GEMM consumes; equals the float32 result cast to bf16, HF:108-114) or float32 (out_f32 = 1: what the processor
 * returns).  Rows in (h / merge, w / merge, merge, merge) order, the frame repeated `temporal` times.  workspace: device scratch of
 * padt_preprocess_workspace_bytes. */
int padt_preprocess_image(const uint8_t* rgb, int H, int W, int out_h, int out_w, const float* mean3_host, const float* std3_host,
                          double rescale_factor, int patch, int temporal, int merge, void* pixel_values, int out_f32, void* workspace,
                          size_t workspace_bytes, padt_stream_t stream);

/* ---- result post-processing of the batched inference harness on the GPU (eval/evaluation_scripts/utils.py:260-264).
 * mask_logits: device bf16, one object's pred_mask plane (row stride in elements), valid region [valid_h, valid_w] = [4H, 4W].
 * Computes F.interpolate(..., size = (out_h, out_w), mode = 'bilinear').sigmoid() > 0.5 with the reference's bf16 rounding and writes
 *   dense_cm   device uint8 [out_w * out_h], COLUMN-major (np.asfortranarray(mask)),
 *   run_starts device int32 [out_h * out_w] capacity: positions in the column-major flattening where the value changes (position 0 is a
 *              start when the first pixel is 1), n_runs device int32 [1] = how many.
 * padt_rle_string (host) turns the starts into pycocotools' compressed RLE "counts" string (maskApi.c rleToString). */
size_t padt_mask_runs_workspace_bytes(int out_h, int out_w);
int padt_mask_to_runs(const void* mask_logits, int64_t row_stride, int valid_h, int valid_w, int out_h, int out_w, uint8_t* dense_cm,
                      int32_t* run_starts, int32_t* n_runs, void* workspace, size_t workspace_bytes, padt_stream_t stream);
int padt_rle_string(const int32_t* run_starts_host, int n_starts, int total, char* out, int cap);

/* Host-only planning helpers of the persistent decode kernel (no device work; CPU tests): shared-memory plan
 * out4 = {slot_bytes, n_slots (always even), activation-region bytes, dynamic shared memory bytes}; stage range of one CTA for an
 * [n_rows, K] matrix: out3 = {K segments per row group, first stage, one-past-last stage}, stage = row_group * nseg + segment. */
int padt_debug_mega_plan(int D, int I, int n_heads, int n_kv_heads, int B, int32_t* out4);
int padt_debug_mega_partition(int n_rows, int K, int n_ctas, int balanced, int cta, int32_t* out3);

#ifdef __cplusplus
}
#endif
#endif /* PADT_B200_H */
