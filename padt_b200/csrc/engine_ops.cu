// Kernel-level C-ABI entry points used by the parity tests (one kernel per call, borrowed device pointers).
#include "engine.h"

extern "C" int padt_op_gemm(int epilogue, const void* A, int lda, const void* W, int ldw, const void* bias, const void* resid, int ldr, void* C,
                            int ldc, const int32_t* out_row_map, int M, int N, int K, padt_stream_t stream) {
    return gemm(epilogue, (const bf16*)A, lda, (const bf16*)W, ldw, (const bf16*)bias, (const bf16*)resid, ldr, (bf16*)C, ldc, out_row_map, M, N, K,
                (cudaStream_t)stream);
}

extern "C" int padt_op_rmsnorm(const void* x, const void* w, void* y, int rows, int dim, float eps, padt_stream_t stream) {
    return launch_rmsnorm((const bf16*)x, dim, (const bf16*)w, (bf16*)y, dim, rows, dim, eps, (cudaStream_t)stream);
}

extern "C" int padt_op_attention(const void* q, const void* k, const void* v, void* o, int64_t q_row, int64_t q_head, int64_t k_row, int64_t k_head,
                                 int64_t v_row, int64_t v_head, int64_t o_row, int64_t o_head, const int32_t* q_off, const int32_t* q_len,
                                 const int32_t* k_off, const int32_t* k_len, const int32_t* k_start, int n_seg, int n_heads, int group,
                                 int max_q_len, int head_dim, int causal, float scale, padt_stream_t stream) {
    AttnParams a = {};
    a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (bf16*)o;
    a.q_row = q_row; a.q_head = q_head; a.k_row = k_row; a.k_head = k_head; a.v_row = v_row; a.v_head = v_head; a.o_row = o_row; a.o_head = o_head;
    a.q_off = q_off; a.q_len = q_len; a.k_off = k_off; a.k_len = k_len; a.k_start = k_start;
    a.n_seg = n_seg; a.n_heads = n_heads; a.group = group; a.max_q_len = max_q_len; a.head_dim = head_dim; a.causal = causal; a.scale = scale;
    // test entry: fetch the segment table so that long segments reach the tcgen05 kernel exactly as the engine's calls do
    std::vector<int> hq(4 * (size_t)(n_seg > 0 ? n_seg : 1));
    AttnSegHost hseg = {hq.data(), hq.data() + n_seg, hq.data() + 2 * n_seg, hq.data() + 3 * n_seg};
    if (n_seg > 0 && n_seg <= 8 && max_q_len >= 512) {
        const int32_t* src[4] = {q_off, q_len, k_off, k_len};
        for (int i = 0; i < 4; ++i)
            PADT_CHECK_CUDA(cudaMemcpyAsync(hq.data() + (size_t)i * n_seg, src[i], (size_t)n_seg * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        PADT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        a.host_seg = &hseg;
    }
    return launch_attention(a, (cudaStream_t)stream);
}

extern "C" int padt_op_attention_split(const void* q, const void* k, const void* v, void* o, int64_t q_row, int64_t q_head, int64_t k_row,
                                       int64_t k_head, int64_t v_row, int64_t v_head, int64_t o_row, int64_t o_head, const int32_t* q_off,
                                       const int32_t* q_len, const int32_t* k_off, const int32_t* k_len, int n_seg, int n_heads, int max_q_len,
                                       int head_dim, float scale, int n_split, float* part_o, float* part_ml, padt_stream_t stream) {
    AttnParams a = {};
    a.q = (const bf16*)q; a.k = (const bf16*)k; a.v = (const bf16*)v; a.o = (bf16*)o;
    a.q_row = q_row; a.q_head = q_head; a.k_row = k_row; a.k_head = k_head; a.v_row = v_row; a.v_head = v_head; a.o_row = o_row; a.o_head = o_head;
    a.q_off = q_off; a.q_len = q_len; a.k_off = k_off; a.k_len = k_len; a.k_start = nullptr;
    a.n_seg = n_seg; a.n_heads = n_heads; a.group = 1; a.max_q_len = max_q_len; a.head_dim = head_dim; a.causal = 0; a.scale = scale;
    a.n_split = n_split; a.part_o = part_o; a.part_ml = part_ml;
    return launch_attention(a, (cudaStream_t)stream);
}

extern "C" int padt_op_rope_neox(void* x, int64_t row_stride, int n_heads, int head_dim, const float* cos, const float* sin, int cs_ld, int rows,
                                 padt_stream_t stream) {
    return launch_rope_neox_f32((bf16*)x, row_stride, n_heads, head_dim, cos, sin, cs_ld, nullptr, rows, (cudaStream_t)stream);
}

// ---- host-only planning helpers of the persistent decode kernel (no device work; used by the CPU tests)
extern "C" int padt_op_logit_mask(void* logits, int64_t ld, int rows_per_sample, int B, int vocab, int n_ext, const int32_t* vrt_lo_host,
                                  const int32_t* vrt_hi_host, padt_stream_t stream) {
    PADT_REQUIRE(logits && vrt_lo_host && vrt_hi_host && B >= 1 && B <= 16 && rows_per_sample >= 1 && n_ext >= vocab && ld >= n_ext, "arguments");
    LogitMaskRanges r = {};
    for (int b = 0; b < B; ++b) {
        PADT_REQUIRE(vrt_lo_host[b] >= 0 && vrt_lo_host[b] <= vrt_hi_host[b] && vrt_hi_host[b] <= n_ext - vocab, "VRT range");
        r.lo[b] = vrt_lo_host[b]; r.hi[b] = vrt_hi_host[b];
    }
    return launch_logit_mask((bf16*)logits, ld, rows_per_sample, B, vocab, n_ext, r, (cudaStream_t)stream);
}

extern "C" int padt_debug_mega_plan(int D, int I, int n_heads, int n_kv_heads, int B, int32_t* out4) {
    PADT_REQUIRE(out4 && n_heads > 0 && D % n_heads == 0, "arguments");
    DecodeParams p = {};
    p.B = B; p.D = D; p.I = I; p.nh = n_heads; p.nkv = n_kv_heads; p.hd = D / n_heads;
    int sb = 0, ns = 0, xb = 0;
    size_t smem = 0;
    if (decode_mega_plan(p, &sb, &ns, &xb, &smem)) { padt_set_error("model does not fit the shared-memory plan"); return -1; }
    out4[0] = sb; out4[1] = ns; out4[2] = xb; out4[3] = (int32_t)smem;
    return 0;
}

extern "C" int padt_debug_mega_partition(int n_rows, int K, int n_ctas, int balanced, int cta, int32_t* out3) {
    PADT_REQUIRE(out3 && n_ctas > 0 && cta >= 0 && cta < n_ctas && K % 16 == 0, "arguments");
    int nseg;
    unsigned int lo, hi;
    decode_mega_stage_range(n_rows, K, n_ctas, balanced != 0, cta, &nseg, &lo, &hi);
    out3[0] = nseg; out3[1] = (int32_t)lo; out3[2] = (int32_t)hi;
    return 0;
}
