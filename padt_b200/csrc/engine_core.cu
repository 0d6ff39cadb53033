// Engine construction: canonical weight order, packed/derived weight arena, handle lifetime.
// Replaces the module construction of PaDTForConditionalGeneration.__init__ (padt.py:115-132) and
// PaDTDecoder.__init__ (padt_decoder.py:131-185) for the inference hot path.
#include "engine.h"

unsigned long long g_padt_launches = 0;
extern "C" unsigned long long padt_launch_count(void) { return g_padt_launches; }
static thread_local std::string g_last_error;
void padt_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* padt_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* padt_version(void) { return "padt_b200 0.1 (sm_100a)"; }

// Must mirror padt_b200/weights.py::weight_shapes (checked name-by-name from Python at create time).
void enumerate_weight_names(const PadtDims& d, std::vector<std::string>& s) {
    s.clear();
    s.push_back("visual.patch_embed.proj.weight");
    for (int i = 0; i < d.vit_depth; ++i) {
        std::string p = "visual.blocks." + std::to_string(i) + ".";
        for (const char* n : {"norm1.weight", "norm2.weight", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                              "mlp.gate_proj.weight", "mlp.gate_proj.bias", "mlp.up_proj.weight", "mlp.up_proj.bias", "mlp.down_proj.weight",
                              "mlp.down_proj.bias"})
            s.push_back(p + n);
    }
    for (const char* n : {"visual.merger.ln_q.weight", "visual.merger.mlp.0.weight", "visual.merger.mlp.0.bias", "visual.merger.mlp.2.weight",
                          "visual.merger.mlp.2.bias", "model.embed_tokens.weight"})
        s.push_back(n);
    for (int i = 0; i < d.n_layers; ++i) {
        std::string p = "model.layers." + std::to_string(i) + ".";
        for (const char* n : {"input_layernorm.weight", "self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight",
                              "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias", "self_attn.o_proj.weight",
                              "post_attention_layernorm.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight"})
            s.push_back(p + n);
    }
    s.push_back("model.norm.weight");
    if (!d.tie_word_embeddings) s.push_back("lm_head.weight");
    if (d.use_proto_proj)
        for (const char* n : {"vis_norm.weight", "vis_norm.bias", "vis_proj.0.weight", "vis_proj.1.weight"}) s.push_back(n);
    std::string q = "vl_decoder.";
    for (const char* n : {"vp_embedding.weight", "bbox_score_mask_tokens.weight", "input_projection.0.weight", "input_projection.1.weight",
                          "input_projection.1.bias", "input_projection.3.weight", "input_projection.3.bias"})
        s.push_back(q + n);
    for (const char* blk : {"low_res_transformer", "high_res_transformer1", "high_res_transformer2"}) {
        std::string p = q + blk + ".";
        for (int n = 1; n <= 6; ++n) s.push_back(p + "norm" + std::to_string(n) + ".weight");
        for (const char* att : {"self_attn", "cross_attn_query_to_image", "cross_attn_image_to_query"})
            for (const char* lin : {"q_proj", "k_proj", "v_proj", "proj"}) {
                s.push_back(p + att + "." + lin + ".weight");
                s.push_back(p + att + "." + lin + ".bias");
            }
        for (const char* n : {"mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias"}) s.push_back(p + n);
    }
    s.push_back(q + "high_res_norm.weight");
    for (const char* head : {"bbox_prediction", "mask_output_mlp"})
        for (const char* n : {".0.weight", ".0.bias", ".2.weight", ".2.bias", ".4.weight", ".4.bias"}) s.push_back(q + head + n);
    for (const char* n : {"score_prediction.weight", "score_prediction.bias", "mask_output_upscaling1.0.weight", "mask_output_upscaling1.0.bias",
                          "mask_output_upscaling1.1.weight", "mask_output_upscaling2.0.weight", "mask_output_upscaling2.0.bias"})
        s.push_back(q + n);
}

static thread_local std::vector<std::string> g_names;
extern "C" int padt_weight_count(const PadtDims* dims) {
    enumerate_weight_names(*dims, g_names);
    return (int)g_names.size();
}
extern "C" const char* padt_weight_name(const PadtDims* dims, int index) {
    enumerate_weight_names(*dims, g_names);
    if (index < 0 || index >= (int)g_names.size()) return "";
    return g_names[index].c_str();
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int engine_build_weights(PadtHandle* h) {
    const PadtDims& d = h->d;
    auto W = [&](const std::string& n) -> const bf16* {
        auto it = h->widx.find(n);
        if (it == h->widx.end()) { padt_set_error("missing weight " + n); return nullptr; }
        return h->w[it->second];
    };
    const int Hv = d.vit_hidden, Iv = d.vit_inter, Ivp = h->vit_ipad;
    const int D = d.D, I = d.I, hd = h->hd;
    const int Hd = d.dec_hidden, Id = d.dec_inter, Idp = h->dec_ipad;
    const int qkv_n = (d.n_heads + 2 * d.n_kv_heads) * hd;
    // ---- arena size
    size_t need = 0;
    auto take = [&](size_t elems) { size_t off = need; need = align_up(need + elems * sizeof(bf16), 256); return off; };
    std::vector<size_t> o_vgu(d.vit_depth), o_vbgu(d.vit_depth), o_vdown(d.vit_depth);
    for (int i = 0; i < d.vit_depth; ++i) { o_vgu[i] = take((size_t)2 * Iv * Hv); o_vbgu[i] = take((size_t)2 * Iv); o_vdown[i] = take((size_t)Hv * Ivp); }
    std::vector<size_t> o_qkv(d.n_layers), o_bqkv(d.n_layers), o_gu(d.n_layers);
    for (int i = 0; i < d.n_layers; ++i) { o_qkv[i] = take((size_t)qkv_n * D); o_bqkv[i] = take(qkv_n); o_gu[i] = take((size_t)2 * I * D); }
    size_t o_dmlp2[3];
    for (int b = 0; b < 3; ++b) o_dmlp2[b] = take((size_t)Hd * Idp);
    // stage images for the persistent decode kernel (decode_mega.cu)
    const int attn_n = d.n_heads * hd;
    std::vector<size_t> o_iqkv(d.n_layers), o_io(d.n_layers), o_igu(d.n_layers), o_idown(d.n_layers);
    for (int i = 0; i < d.n_layers; ++i) {
        o_iqkv[i] = take(decode_mega_image_elems(qkv_n, D));
        o_io[i] = take(decode_mega_image_elems(D, attn_n));
        o_igu[i] = take(decode_mega_image_elems(2 * I, D));
        o_idown[i] = take(decode_mega_image_elems(D, I));
    }
    size_t o_ihead = take(decode_mega_image_elems(d.vocab, D));
    RC(h->arena.ensure(need));
    PADT_CHECK_CUDA(cudaMemset(h->arena.p, 0, need));
    uint8_t* base = h->arena.as<uint8_t>();
    auto at = [&](size_t off) { return reinterpret_cast<bf16*>(base + off); };
    cudaStream_t st = 0;

    // ---- ViT
    h->w_patch = W("visual.patch_embed.proj.weight");
    h->vit.resize(d.vit_depth);
    for (int i = 0; i < d.vit_depth; ++i) {
        std::string p = "visual.blocks." + std::to_string(i) + ".";
        VitBlockW& b = h->vit[i];
        b.norm1 = W(p + "norm1.weight"); b.norm2 = W(p + "norm2.weight");
        b.wqkv = W(p + "attn.qkv.weight"); b.bqkv = W(p + "attn.qkv.bias");
        b.wproj = W(p + "attn.proj.weight"); b.bproj = W(p + "attn.proj.bias");
        RC(launch_interleave_rows(W(p + "mlp.gate_proj.weight"), W(p + "mlp.up_proj.weight"), at(o_vgu[i]), Iv, Hv, st));
        RC(launch_interleave_rows(W(p + "mlp.gate_proj.bias"), W(p + "mlp.up_proj.bias"), at(o_vbgu[i]), Iv, 1, st));
        PADT_CHECK_CUDA(cudaMemcpy2DAsync(at(o_vdown[i]), (size_t)Ivp * 2, W(p + "mlp.down_proj.weight"), (size_t)Iv * 2, (size_t)Iv * 2, Hv,
                                          cudaMemcpyDeviceToDevice, st));
        b.wgu = at(o_vgu[i]); b.bgu = at(o_vbgu[i]); b.wdown = at(o_vdown[i]); b.bdown = W(p + "mlp.down_proj.bias");
    }
    h->mg_ln = W("visual.merger.ln_q.weight");
    h->mg_w0 = W("visual.merger.mlp.0.weight"); h->mg_b0 = W("visual.merger.mlp.0.bias");
    h->mg_w2 = W("visual.merger.mlp.2.weight"); h->mg_b2 = W("visual.merger.mlp.2.bias");
    // ---- prototypes
    if (d.use_proto_proj) {
        h->vn_w = W("vis_norm.weight"); h->vn_b = W("vis_norm.bias"); h->vp_w0 = W("vis_proj.0.weight"); h->vp_w1 = W("vis_proj.1.weight");
    }
    // ---- LLM
    h->llm.resize(d.n_layers);
    for (int i = 0; i < d.n_layers; ++i) {
        std::string p = "model.layers." + std::to_string(i) + ".";
        DecodeLayerW& l = h->llm[i];
        l.ln1 = W(p + "input_layernorm.weight"); l.ln2 = W(p + "post_attention_layernorm.weight");
        bf16* qkv = at(o_qkv[i]);
        bf16* bq = at(o_bqkv[i]);
        size_t nq = (size_t)d.n_heads * hd, nk = (size_t)d.n_kv_heads * hd;
        PADT_CHECK_CUDA(cudaMemcpyAsync(qkv, W(p + "self_attn.q_proj.weight"), nq * D * 2, cudaMemcpyDeviceToDevice, st));
        PADT_CHECK_CUDA(cudaMemcpyAsync(qkv + nq * D, W(p + "self_attn.k_proj.weight"), nk * D * 2, cudaMemcpyDeviceToDevice, st));
        PADT_CHECK_CUDA(cudaMemcpyAsync(qkv + (nq + nk) * D, W(p + "self_attn.v_proj.weight"), nk * D * 2, cudaMemcpyDeviceToDevice, st));
        PADT_CHECK_CUDA(cudaMemcpyAsync(bq, W(p + "self_attn.q_proj.bias"), nq * 2, cudaMemcpyDeviceToDevice, st));
        PADT_CHECK_CUDA(cudaMemcpyAsync(bq + nq, W(p + "self_attn.k_proj.bias"), nk * 2, cudaMemcpyDeviceToDevice, st));
        PADT_CHECK_CUDA(cudaMemcpyAsync(bq + nq + nk, W(p + "self_attn.v_proj.bias"), nk * 2, cudaMemcpyDeviceToDevice, st));
        l.wqkv = qkv; l.bqkv = bq;
        l.wo = W(p + "self_attn.o_proj.weight");
        RC(launch_interleave_rows(W(p + "mlp.gate_proj.weight"), W(p + "mlp.up_proj.weight"), at(o_gu[i]), I, D, st));
        l.wgu = at(o_gu[i]);
        l.wdown = W(p + "mlp.down_proj.weight");
        RC(launch_pack_stage_image(l.wqkv, qkv_n, D, at(o_iqkv[i]), st));
        RC(launch_pack_stage_image(l.wo, D, attn_n, at(o_io[i]), st));
        RC(launch_pack_stage_image(l.wgu, 2 * I, D, at(o_igu[i]), st));
        RC(launch_pack_stage_image(l.wdown, D, I, at(o_idown[i]), st));
        l.iqkv = at(o_iqkv[i]); l.io = at(o_io[i]); l.igu = at(o_igu[i]); l.idown = at(o_idown[i]);
    }
    h->final_norm = W("model.norm.weight");
    h->embed = W("model.embed_tokens.weight");
    h->head = d.tie_word_embeddings ? h->embed : W("lm_head.weight");
    RC(launch_pack_stage_image(h->head, d.vocab, D, at(o_ihead), st));
    h->head_img = at(o_ihead);
    // ---- PaDT decoder
    std::string q = "vl_decoder.";
    h->dec_vp_emb = W(q + "vp_embedding.weight"); h->dec_tokens3 = W(q + "bbox_score_mask_tokens.weight");
    h->ip_ln = W(q + "input_projection.0.weight"); h->ip_w1 = W(q + "input_projection.1.weight"); h->ip_b1 = W(q + "input_projection.1.bias");
    h->ip_w3 = W(q + "input_projection.3.weight"); h->ip_b3 = W(q + "input_projection.3.bias");
    h->dec_hr_norm = W(q + "high_res_norm.weight");
    const char* blks[3] = {"low_res_transformer", "high_res_transformer1", "high_res_transformer2"};
    for (int b = 0; b < 3; ++b) {
        std::string p = q + blks[b] + ".";
        DecBlockW& k = h->dec_blocks[b];
        for (int n = 0; n < 6; ++n) k.norm[n] = W(p + "norm" + std::to_string(n + 1) + ".weight");
        auto att = [&](const std::string& a, DecAttnW& o) {
            o.wq = W(p + a + ".q_proj.weight"); o.bq = W(p + a + ".q_proj.bias");
            o.wk = W(p + a + ".k_proj.weight"); o.bk = W(p + a + ".k_proj.bias");
            o.wv = W(p + a + ".v_proj.weight"); o.bv = W(p + a + ".v_proj.bias");
            o.wp = W(p + a + ".proj.weight"); o.bp = W(p + a + ".proj.bias");
        };
        att("self_attn", k.self_attn); att("cross_attn_query_to_image", k.q2i); att("cross_attn_image_to_query", k.i2q);
        k.w0 = W(p + "mlp.0.weight"); k.b0 = W(p + "mlp.0.bias"); k.b2 = W(p + "mlp.2.bias");
        PADT_CHECK_CUDA(cudaMemcpy2DAsync(at(o_dmlp2[b]), (size_t)Idp * 2, W(p + "mlp.2.weight"), (size_t)Id * 2, (size_t)Id * 2, Hd,
                                          cudaMemcpyDeviceToDevice, st));
        k.w2 = at(o_dmlp2[b]);
    }
    auto mlp3 = [&](const std::string& n, Mlp3W& m) {
        m.w0 = W(q + n + ".0.weight"); m.b0 = W(q + n + ".0.bias"); m.w2 = W(q + n + ".2.weight"); m.b2 = W(q + n + ".2.bias");
        m.w4 = W(q + n + ".4.weight"); m.b4 = W(q + n + ".4.bias");
    };
    mlp3("bbox_prediction", h->bbox_mlp); mlp3("mask_output_mlp", h->mask_mlp);
    h->score_w = W(q + "score_prediction.weight"); h->score_b = W(q + "score_prediction.bias");
    h->up1_w = W(q + "mask_output_upscaling1.0.weight"); h->up1_b = W(q + "mask_output_upscaling1.0.bias");
    h->up1_ln = W(q + "mask_output_upscaling1.1.weight");
    h->up2_w = W(q + "mask_output_upscaling2.0.weight"); h->up2_b = W(q + "mask_output_upscaling2.0.bias");
    PADT_CHECK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int padt_create(const PadtDims* dims, const void* const* weights, int n_weights, const void* llm_cos, const void* llm_sin, int max_pos,
                           PadtHandle** out) {
    PADT_REQUIRE(dims && weights && out, "null argument");
    std::vector<std::string> names;
    enumerate_weight_names(*dims, names);
    PADT_REQUIRE((int)names.size() == n_weights, "weight count does not match the canonical inventory");
    const PadtDims& d = *dims;
    PADT_REQUIRE(d.vit_hidden % d.vit_heads == 0 && d.D % d.n_heads == 0 && d.dec_hidden % d.dec_heads == 0, "head split");
    PADT_REQUIRE(d.vit_hidden / d.vit_heads == 80 && d.dec_hidden / d.dec_heads == 80, "ViT / PaDT-decoder kernels are built for head_dim 80");
    PADT_REQUIRE(d.D / d.n_heads == 128, "LLM kernels are built for head_dim 128");
    PADT_REQUIRE(d.vit_hidden % 8 == 0 && d.D % 8 == 0 && d.I % 8 == 0 && d.dec_hidden % 16 == 0, "dims must keep 16-byte rows");
    PADT_REQUIRE(d.vit_merge == 2, "spatial_merge_size 2 (the decoder hard-codes 4 sub-patches, padt.py:365)");
    PADT_REQUIRE((d.mrope0 + d.mrope1 + d.mrope2) * 2 == 128, "mrope sections");
    PADT_REQUIRE(d.dec_hidden == d.vit_hidden, "padt_decoder.py:220 adds decoder-width and ViT-width features");
    PADT_REQUIRE(d.vit_n_fullatt <= 8, "fullatt list");
    for (int i = 0; i < n_weights; ++i) PADT_REQUIRE(weights[i] != nullptr && (((uintptr_t)weights[i]) & 15) == 0, "weight pointers must be 16-byte aligned");
    PadtHandle* h = new PadtHandle();
    h->d = d;
    h->vit_hd = 80; h->dec_hd = 80; h->hd = 128;
    h->vit_ipad = (d.vit_inter + 7) / 8 * 8;
    h->dec_ipad = (d.dec_inter + 7) / 8 * 8;
    h->w.resize(n_weights);
    for (int i = 0; i < n_weights; ++i) { h->w[i] = reinterpret_cast<const bf16*>(weights[i]); h->widx[names[i]] = i; }
    h->cos_tab = reinterpret_cast<const bf16*>(llm_cos);
    h->sin_tab = reinterpret_cast<const bf16*>(llm_sin);
    h->max_pos = max_pos;
    int rc = engine_build_weights(h);
    if (rc) { delete h; return rc; }
    {   // device copy of the per-layer pointer table + grid-barrier word for the persistent decode kernel
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, dev);
        if (h->llm_dev.ensure(h->llm.size() * sizeof(DecodeLayerW)) || h->gbar.ensure(256)) { delete h; return -2; }
        cudaMemcpy(h->llm_dev.p, h->llm.data(), h->llm.size() * sizeof(DecodeLayerW), cudaMemcpyHostToDevice);
        cudaMemset(h->gbar.p, 0, 256);
    }
    *out = h;
    return 0;
}

extern "C" int padt_set_rope_table(PadtHandle* h, const void* llm_cos, const void* llm_sin, int max_pos) {
    PADT_REQUIRE(h && llm_cos && llm_sin && max_pos > 0, "null argument");
    cudaDeviceSynchronize();   // a decode launch in flight may still read the old table
    h->cos_tab = reinterpret_cast<const bf16*>(llm_cos);
    h->sin_tab = reinterpret_cast<const bf16*>(llm_sin);
    h->max_pos = max_pos;
    h->gen_ready = false;      // generation state (and any captured graph) refers to the old table
    if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
    return 0;
}

extern "C" int padt_destroy(PadtHandle* h) {
    if (!h) return 0;
    cudaDeviceSynchronize();
    if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
    if (h->own_stream) { cudaStreamDestroy(h->own_stream); cudaEventDestroy(h->ev_in); cudaEventDestroy(h->ev_out); }
    for (auto& kv : h->vit_plans) { kv.second->ints.release(); delete kv.second; }
    for (DevBuf* b : {&h->arena, &h->vit_ws, &h->proto_ws, &h->llm_ws, &h->kv, &h->dec_ws, &h->state, &h->seg_ints, &h->eos_dev, &h->dec_dws,
                      &h->dec_ints, &h->llm_dev, &h->gbar, &h->stage_tab, &h->stage_off})
        b->release();
    delete h;
    return 0;
}

extern "C" int padt_set_option(PadtHandle* h, const char* name, int value) {
    PADT_REQUIRE(h && name, "null");
    std::string n(name);
    if (n == "use_graph") { h->use_graph = value; return 0; }
    if (n == "gemm_tc") { g_padt_gemm_tc = value; return 0; }
    if (n == "decode_impl") {
        if (h->decode_impl != value) {
            // the persistent kernel's split-KV and shared-row-group counters are epochs counted in ITS launches since the prefill:
            // switching implementation in the middle of a generation would wait for epochs that never come.  Make the next
            // padt_decode fail loudly ("padt_prefill must run first") instead.
            h->gen_ready = false;
        }
        h->decode_impl = value;
        if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
        return 0;
    }
    padt_set_error("unknown option " + n);
    return -1;
}
