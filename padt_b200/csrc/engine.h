// Host-side engine state shared by the engine_*.cu translation units.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "kernels.h"
#include "../../include/padt_b200.h"

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = need + (need >> 3);
        PADT_CHECK_CUDA(cudaMalloc(&p, want));
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct VitBlockW {
    const bf16 *norm1, *norm2, *wqkv, *bqkv, *wproj, *bproj, *wgu, *bgu, *wdown, *bdown;
};
struct DecAttnW { const bf16 *wq, *bq, *wk, *bk, *wv, *bv, *wp, *bp; };
struct DecBlockW {
    const bf16* norm[6];
    DecAttnW self_attn, q2i, i2q;
    const bf16 *w0, *b0, *w2, *b2;   // mlp (w2 repacked with padded K)
};
struct Mlp3W { const bf16 *w0, *b0, *w2, *b2, *w4, *b4; };

// cached per-grid-signature ViT index tables (device)
struct VitPlan {
    int P = 0, M = 0, n_win = 0, n_full = 0, max_win = 0, max_full = 0;
    DevBuf ints;   // all int32 tables packed
    int *embed_row_map, *hpos, *wpos, *merge_row_map;
    int *win_off, *win_len, *full_off, *full_len;
    std::vector<int> full_off_host, full_len_host;   // host copies: the tcgen05 attention builds its tensor maps per segment
};

#define PADT_VIT_PLAN_CACHE 16

struct PadtHandle {
    PadtDims d;
    int vit_hd, dec_hd, hd;          // head dims
    int vit_ipad, dec_ipad;          // intermediate sizes padded to 8
    std::vector<const bf16*> w;      // raw weight pointers in canonical order
    std::map<std::string, int> widx;
    DevBuf arena;                    // packed / derived weights
    // ViT
    const bf16* w_patch;
    std::vector<VitBlockW> vit;
    const bf16 *mg_ln, *mg_w0, *mg_b0, *mg_w2, *mg_b2;
    std::map<std::string, VitPlan*> vit_plans;      // bounded LRU (PADT_VIT_PLAN_CACHE entries), keyed by the batch's grid signature
    std::vector<std::string> vit_plan_order;        // least recently used first
    DevBuf vit_ws;
    // prototypes
    const bf16 *vn_w, *vn_b, *vp_w0, *vp_w1;
    DevBuf proto_ws;
    // LLM
    std::vector<DecodeLayerW> llm;
    const bf16 *final_norm, *head, *embed, *head_img;
    const bf16 *cos_tab, *sin_tab;
    int max_pos;
    DevBuf llm_ws, kv, dec_ws, state, seg_ints;
    // generation state
    DecodeParams dp;
    PadtGenBuffers gen;
    int gen_B = 0, gen_L = 0, s_max = 0;
    bool gen_ready = false;
    int use_graph = 1;
    int stage_tab_split = -1;        // partition mode the cached stage table was built for
    int decode_impl = 1;             // 1 = persistent kernel (decode_mega.cu) when it fits, 0 = per-layer kernels (decode.cu)
    int n_sms = 0;
    DevBuf llm_dev, gbar, stage_tab, stage_off;
    bool stage_tab_ready = false;
    cudaGraphExec_t step_graph = nullptr;
    int graph_kernels = 0;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    const void* graph_key_protos = nullptr;
    int graph_key_B = 0;
    DevBuf eos_dev;
    // decoder
    const bf16 *dec_vp_emb, *dec_tokens3, *ip_ln, *ip_w1, *ip_b1, *ip_w3, *ip_b3, *dec_hr_norm;
    DecBlockW dec_blocks[3];
    Mlp3W bbox_mlp, mask_mlp;
    const bf16 *score_w, *score_b, *up1_w, *up1_b, *up1_ln, *up2_w, *up2_b;
    DevBuf dec_dws, dec_ints;
    std::vector<int> dec_ints_host;   // what dec_ints holds: a repeated object layout skips the upload and its stream sync
};

int engine_build_weights(PadtHandle* h);
void enumerate_weight_names(const PadtDims& d, std::vector<std::string>& names);
int launch_gelu(bf16* x, long long n, cudaStream_t st);
int launch_mask_head(const bf16* sub, const bf16* hyper, const int* cu_patch, const int* obj_w, int n_obj, int n_patches, int ch, int hmax,
                     int wmax, bf16* masks, cudaStream_t st);
int launch_build_queries(const bf16* tokens3, const bf16* vp_proj, const bf16* vp_emb, const int* src, int rows, int dim, bf16* out,
                         cudaStream_t st);
int launch_add_rmsnorm_rep_map(const bf16* low, int rep, const bf16* high, const int* high_row_map, const bf16* w, bf16* y, int rows, int dim,
                               float eps, cudaStream_t st);
int launch_interleave_rows(const bf16* a, const bf16* b, bf16* out, int rows, int cols, cudaStream_t st);

inline int gemm(int epi, const bf16* A, int lda, const bf16* W, int ldw, const bf16* bias, const bf16* resid, int ldr, bf16* C, int ldc,
                const int* map, int M, int N, int K, cudaStream_t st) {
    GemmParams p;
    p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.bias = bias; p.resid = resid; p.ldr = ldr; p.C = C; p.ldc = ldc;
    p.out_row_map = map; p.M = M; p.N = N; p.K = K;
    return launch_gemm(epi, p, st);
}
#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
