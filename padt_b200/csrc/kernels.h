// Internal launcher interface between the host engine (engine.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <vector>

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ dense projections (gemm.cu)
enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_SWIGLU = 3 };
struct GemmParams {
    const bf16* A; int lda;      // [M,K] activations
    const bf16* W; int ldw;      // [N,K] nn.Linear weight
    const bf16* bias;            // [N] or null
    const bf16* resid; int ldr;  // EPI_RESID: C = resid + (A W^T + b)
    bf16* C; int ldc;
    const int* out_row_map;      // optional: output row = map[m] (window permutation fused into the store)
    int M, N, K;
    int dbg_skip_epilogue;       // measurement only (PADT_GEMM_NOEPI=1): accumulators are drained but nothing is computed / stored
};
int launch_gemm(int epi, const GemmParams& p, cudaStream_t st);
// tcgen05 + TMA path for the large shapes (gemm_tc.cu); launch_gemm dispatches to it when eligible and enabled
bool gemm_tc_eligible(const GemmParams& p);
int launch_gemm_tc(int epi, const GemmParams& p, cudaStream_t st);
extern int g_padt_gemm_tc;   // 1 = persistent tcgen05 kernel when eligible (default), 2 = one-tile tcgen05 kernel, 0 = mma.sync only

// ------------------------------------------------------------------ segmented flash attention (attention.cu)
struct AttnParams {
    const bf16 *q, *k, *v;
    bf16* o;
    long long q_row, q_head, k_row, k_head, v_row, v_head, o_row, o_head;  // element strides
    const int *q_off, *q_len, *k_off, *k_len, *k_start;                    // device int32 [n_seg] (k_start optional)
    int n_seg, n_heads, group, max_q_len, head_dim, causal;
    float scale;
    // split-KV (small-q / large-kv cross attention of the PaDT decoder): n_split > 1 -> the key range of every (segment, head) is cut into
    // n_split parts, each CTA writes an un-normalised partial (fp32 o[64][head_dim], m, l) and a merge kernel combines them.  Needs
    // max_q_len <= 64, non-causal, and the two workspaces: part_o [n_seg*n_heads*n_split*64*head_dim], part_ml [n_seg*n_heads*n_split*64*2].
    int n_split;
    float *part_o, *part_ml;
    // host copies of q_off / q_len / k_off / k_len (optional): with them the long-segment calls take the tcgen05 + TMA kernel
    // (attention_tc.cu), whose tensor maps are built per segment on the host
    const struct AttnSegHost* host_seg;
};
struct AttnSegHost { const int *q_off, *q_len, *k_off, *k_len; };
int launch_attention(const AttnParams& p, cudaStream_t st);
bool attention_tc_eligible(const AttnParams& p, const AttnSegHost* hs);
int launch_attention_tc(const AttnParams& p, const AttnSegHost& hs, cudaStream_t st);

// ------------------------------------------------------------------ row-wise / elementwise (elementwise.cu)
int launch_rmsnorm(const bf16* x, int ldx, const bf16* w, bf16* y, int ldy, int rows, int dim, float eps, cudaStream_t st);
int launch_layernorm(const bf16* x, const bf16* w, const bf16* b, bf16* y, int rows, int dim, float eps, cudaStream_t st);
int launch_add(const bf16* a, const bf16* b, bf16* y, long long n, cudaStream_t st);
int launch_add_bcast_row(const bf16* a, const bf16* row, bf16* y, int rows, int dim, cudaStream_t st);
int launch_gather_rows(const bf16* src, int lds, const int* idx, bf16* dst, int ldd, int rows, int dim, cudaStream_t st);
int launch_copy_rows_strided(const bf16* src, long long lds, bf16* dst, long long ldd, int rows, int dim, cudaStream_t st);
// ViT / decoder rotary: x[row, head, :hd] rotated in fp32 with cos/sin rows of width `cs_ld` (first hd/2 columns used)
int launch_rope_neox_f32(bf16* x, long long row_stride, int n_heads, int hd, const float* cos, const float* sin, int cs_ld,
                         const int* cs_row_map, int rows, cudaStream_t st);
// ViT cos/sin table build: out[p, 0:hd] from per-position tables (hpos/wpos in merge-block order, window permuted)
int launch_vit_cos_sin(const float* tab_cos, const float* tab_sin, int quarter, const int* hpos, const int* wpos,
                       float* cos, float* sin, int rows, cudaStream_t st);
// LLM M-RoPE in bf16 (HF:659-668) + KV-cache write.  qkv [T, (nh+2nkv)*hd]; pos3 [3,T] int32; tables [max_pos, hd/2] bf16
int launch_mrope_kv(bf16* qkv, int n_tok, int nh, int nkv, int hd, const int* pos3, const bf16* cos_tab, const bf16* sin_tab,
                    int max_pos, int sec0, int sec1, bf16* kcache, bf16* vcache, int B, int L, int s_max, cudaStream_t st);
// embedding gather from [embed ; prototypes] and image-feature scatter (padt.py:193-219)
int launch_embed_gather(const long long* ids, int n_tok, const bf16* embed, int vocab, const bf16* protos, int n_protos,
                        const bf16* image_embeds, const int* image_slot, int dim, bf16* out, int* err_flag, cudaStream_t st);

int launch_sigmoid(bf16* x, long long n, cudaStream_t st);
struct LogitMaskRanges { int lo[16], hi[16]; };
int launch_logit_mask(bf16* logits, long long ld, int rows_per_sample, int B, int vocab, int n_ext, const LogitMaskRanges& r, cudaStream_t st);

// ------------------------------------------------------------------ decode step (decode.cu)
struct DecodeLayerW {
    const bf16 *ln1, *wqkv, *bqkv, *wo, *ln2, *wgu, *wdown;
    // "stage images" of the same matrices for the persistent decode kernel: [row group of 8][K segment][8 rows][Kseg + 8 pad],
    // i.e. exactly the padded shared-memory picture of one ring stage, so a stage is ONE contiguous TMA bulk copy
    const bf16 *iqkv, *io, *igu, *idown;
};
// Device state block (ints): [0]=tokens sampled so far, [1]=rows still unfinished, [2]=padded prompt length L, [3]=reserved,
// then five per-row arrays of PADT_STATE_BMAX ints: kv_start, rope_delta, vrt_lo, vrt_hi, unfinished.
#define PADT_STATE_HDR 4
#define PADT_STATE_BMAX 16
#define PADT_STATE_INTS (PADT_STATE_HDR + 5 * PADT_STATE_BMAX)
struct DecodeParams {
    int B, D, I, nh, nkv, hd, n_layers, vocab, n_protos, s_max;
    float eps;
    const DecodeLayerW* layers;   // host array [n_layers]
    const bf16 *final_norm, *head, *embed, *protos;
    const bf16 *head_img;          // stage image of the head rows (persistent kernel)
    int nseg_d, nseg_a, nseg_i;    // K segments per row for K = D, nh*hd, I (fixed at pack time)
    const bf16 *cos_tab, *sin_tab; int max_pos;
    bf16 *kcache, *vcache;        // [layers][B][nkv][s_max][hd]
    bf16 *h, *qkv, *act;          // [B,D], [B,(nh+2nkv)hd], [B,max(I, nh*hd)]
    float *part_o, *part_ml; int* attn_counters; int n_split;
    unsigned long long* xpart;           // persistent kernel, balanced partition: [n_sms + 1][8][8] words {epoch, fp32 partial} of shared row groups
    float *amax_val; int *amax_idx; int n_amax_blocks;
    bf16 *last_hidden;            // [B,D] final-norm hidden of the latest forward
    bf16 *hidden_steps;           // optional [max_steps,B,D]; entry k is the state that predicted token k
    long long *tokens_out;        // [max_steps,B] tokens appended to the sequences
    long long *argmax_out;        // [max_steps,B] raw arg-max of every step (differs from tokens_out when forced / finished)
    const long long *forced;      // optional [max_steps,B] teacher-forced tokens
    int max_steps;
    bf16 *logits_out; int logits_ld;   // optional [B, logits_ld]: masked bf16 logits of the latest forward (sampling / logits processors)
    int *dev_state;
    long long *cur_token;         // [B]
    const long long *eos_ids; int n_eos; long long pad_id;
};
int decode_logits_blocks(int n_rows);
int launch_sample_embed(const DecodeParams& p, cudaStream_t st);
int launch_final_logits(const DecodeParams& p, cudaStream_t st);
int launch_decode_layers(const DecodeParams& p, cudaStream_t st);
// persistent one-launch-per-token kernel (decode_mega.cu)
int decode_mega_plan(const DecodeParams& p, int* slot_bytes, int* n_slots, int* x_bytes, size_t* smem_bytes);
bool decode_mega_can_run(const DecodeParams& p, int n_sms);
// stage-image packing (host side of the persistent kernel's weight layout)
int decode_mega_nseg(int K);
size_t decode_mega_image_elems(int n_rows, int K);
int launch_pack_stage_image(const bf16* W, int n_rows, int K, bf16* out, cudaStream_t st);
int decode_mega_build_table(const DecodeParams& p, const DecodeLayerW* layers_host, int G, bool split, std::vector<uint64_t>& tab, std::vector<int>& off);
void decode_mega_stage_range(int n_rows, int K, int G, bool balanced, int c, int* nseg, unsigned int* lo, unsigned int* hi);
bool decode_mega_split();   // balanced (stage-granular) partition enabled (PADT_MEGA_FLAGS bit 128 / default)
int launch_decode_mega(const DecodeParams& p, const DecodeLayerW* layers_dev, const void* stage_tab, const int* stage_off, unsigned int* gbar,
                       int n_sms, cudaStream_t st);
