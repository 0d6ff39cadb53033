// LLM prefill and the decode loop (host orchestration).
// Reference: forward_main padt.py:147-340 (prefill branch :183-219, decode branch :221-229, M-RoPE positions :256-277,
// extended logits :292-301) and the greedy loop of _sample_vision_token padt.py:671-757.
#include "engine.h"

namespace {

int setup_decode_params(PadtHandle* h, int B, int L, int n_protos, const void* prototypes) {
    const PadtDims& d = h->d;
    DecodeParams& p = h->dp;
    p.B = B; p.D = d.D; p.I = d.I; p.nh = d.n_heads; p.nkv = d.n_kv_heads; p.hd = h->hd; p.n_layers = d.n_layers;
    p.vocab = d.vocab; p.n_protos = n_protos; p.s_max = h->s_max; p.eps = d.rms_eps;
    p.layers = h->llm.data();
    p.final_norm = h->final_norm; p.head = h->head; p.embed = h->embed; p.protos = (const bf16*)prototypes;
    p.head_img = h->head_img;
    p.nseg_d = decode_mega_nseg(d.D); p.nseg_a = decode_mega_nseg(d.n_heads * h->hd); p.nseg_i = decode_mega_nseg(d.I);
    p.cos_tab = h->cos_tab; p.sin_tab = h->sin_tab; p.max_pos = h->max_pos;
    {   // split-KV factor: enough (row, kv head, split) items to cover the SMs, capped at 64 partials per head
        int ns = h->n_sms / (B * d.n_kv_heads);
        p.n_split = ns < 1 ? 1 : (ns > 64 ? 64 : ns);
    }
    p.n_amax_blocks = decode_logits_blocks(d.vocab + n_protos);
    if (p.n_amax_blocks < h->n_sms) p.n_amax_blocks = h->n_sms;
    // decode workspace: h [B,D] | qkv | act | part_o | part_ml | amax_val | amax_idx | counters | cur_token
    const int qkv_n = (d.n_heads + 2 * d.n_kv_heads) * h->hd;
    const int act_n = d.I > d.n_heads * h->hd ? d.I : d.n_heads * h->hd;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_h = take((size_t)B * d.D * 2), o_qkv = take((size_t)B * qkv_n * 2), o_act = take((size_t)B * act_n * 2);
    size_t o_po = take((size_t)B * d.n_heads * p.n_split * h->hd * 4), o_pml = take((size_t)B * d.n_heads * p.n_split * 2 * 4);
    size_t o_av = take((size_t)p.n_amax_blocks * B * 4), o_ai = take((size_t)p.n_amax_blocks * B * 4);
    size_t o_cnt = take((size_t)B * d.n_kv_heads * 4 * 2), o_tok = take((size_t)B * 8);
    size_t o_xp = take((size_t)(h->n_sms + 1) * 64 * 8);   // [B*nkv] re-armed counters (decode.cu) + [B*nkv] monotonic (decode_mega.cu)
    bool fresh = off > h->dec_ws.bytes;
    RC(h->dec_ws.ensure(off));
    uint8_t* b = h->dec_ws.as<uint8_t>();
    p.h = (bf16*)(b + o_h); p.qkv = (bf16*)(b + o_qkv); p.act = (bf16*)(b + o_act);
    p.part_o = (float*)(b + o_po); p.part_ml = (float*)(b + o_pml);
    p.amax_val = (float*)(b + o_av); p.amax_idx = (int*)(b + o_ai);
    p.attn_counters = (int*)(b + o_cnt); p.cur_token = (long long*)(b + o_tok);
    p.xpart = (unsigned long long*)(b + o_xp);
    PADT_CHECK_CUDA(cudaMemset(p.xpart, 0, (size_t)(h->n_sms + 1) * 64 * 8));   // epoch 0 = nothing published
    (void)fresh;
    PADT_CHECK_CUDA(cudaMemset(p.attn_counters, 0, (size_t)B * d.n_kv_heads * 4 * 2));
    const size_t layer_kv = (size_t)B * d.n_kv_heads * h->s_max * h->hd;
    p.kcache = h->kv.as<bf16>();
    p.vcache = p.kcache + layer_kv * d.n_layers;
    RC(h->state.ensure(PADT_STATE_INTS * sizeof(int)));
    p.dev_state = h->state.as<int>();
    p.last_hidden = (bf16*)h->gen.last_hidden; p.hidden_steps = (bf16*)h->gen.hidden_steps;
    p.tokens_out = (long long*)h->gen.tokens_out; p.argmax_out = (long long*)h->gen.argmax_out;
    p.forced = (const long long*)h->gen.forced; p.max_steps = h->gen.max_steps;
    p.logits_out = (bf16*)h->gen.logits_out; p.logits_ld = h->gen.logits_ld;
    return 0;
}

int run_step(PadtHandle* h, cudaStream_t st) {
    if (h->decode_impl == 1) {
        if (decode_mega_can_run(h->dp, h->n_sms)) {
            const bool split = decode_mega_split();
            if (!h->stage_tab_ready || h->stage_tab_split != (int)split) {   // the weight schedule only depends on the partition mode: flatten it once
                std::vector<uint64_t> tab;
                std::vector<int> off;
                RC(decode_mega_build_table(h->dp, h->llm.data(), h->n_sms, split, tab, off));
                h->stage_tab_split = (int)split;
                PADT_CHECK_CUDA(cudaStreamSynchronize(st));   // a previous launch may still read the old table
                RC(h->stage_tab.ensure(tab.size() * 8 + 16));
                RC(h->stage_off.ensure(off.size() * 4));
                PADT_CHECK_CUDA(cudaMemcpy(h->stage_tab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice));
                PADT_CHECK_CUDA(cudaMemcpy(h->stage_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
                h->stage_tab_ready = true;
            }
            return launch_decode_mega(h->dp, h->llm_dev.as<DecodeLayerW>(), h->stage_tab.p, h->stage_off.as<int>(), h->gbar.as<unsigned int>(),
                                      h->n_sms, st);
        }
    }
    RC(launch_sample_embed(h->dp, st));
    RC(launch_decode_layers(h->dp, st));
    RC(launch_final_logits(h->dp, st));
    return 0;
}

}  // namespace

extern "C" int padt_prefill(PadtHandle* h, const int64_t* input_ids, const int32_t* image_slot, const int32_t* pos3,
                            const int32_t* kv_start_host, const int32_t* rope_delta_host, const int32_t* vrt_lo_host, const int32_t* vrt_hi_host,
                            const void* image_embeds, const void* prototypes, int n_protos, int B, int L, const int64_t* eos_ids_host, int n_eos,
                            int64_t pad_id, const PadtGenBuffers* bufs, padt_stream_t stream) {
    PADT_REQUIRE(h && input_ids && pos3 && kv_start_host && rope_delta_host && vrt_lo_host && vrt_hi_host && bufs, "null argument");
    PADT_REQUIRE(B >= 1 && B <= 8 && L >= 1, "batch must be 1..8 rows per GPU");
    PADT_REQUIRE(bufs->tokens_out && bufs->last_hidden && bufs->max_steps >= 1, "generation buffers");
    PADT_REQUIRE(prototypes != nullptr || n_protos == 0, "prototypes");
    PADT_REQUIRE(bufs->logits_out == nullptr || bufs->logits_ld >= h->d.vocab + n_protos, "logits_out row length");
    cudaStream_t st = (cudaStream_t)stream;
    const PadtDims& d = h->d;
    const int D = d.D, I = d.I, nh = d.n_heads, nkv = d.n_kv_heads, hd = h->hd;
    const int T = B * L, qkv_n = (nh + 2 * nkv) * hd;
    h->gen = *bufs;
    h->gen_B = B; h->gen_L = L;
    h->gen_ready = false;
    // a graph captured for an earlier request bakes in that request's pointers
    if (h->step_graph) { cudaGraphExecDestroy(h->step_graph); h->step_graph = nullptr; }
    const int s_need = L + bufs->max_steps + 1;
    h->s_max = (s_need + 63) / 64 * 64;
    const size_t layer_kv = (size_t)B * nkv * h->s_max * hd;
    RC(h->kv.ensure(layer_kv * d.n_layers * 2 * sizeof(bf16)));
    RC(setup_decode_params(h, B, L, n_protos, prototypes));
    DecodeParams& p = h->dp;
    p.n_eos = n_eos; p.pad_id = pad_id;
    RC(h->eos_dev.ensure((size_t)(n_eos > 0 ? n_eos : 1) * 8));
    if (n_eos > 0) PADT_CHECK_CUDA(cudaMemcpyAsync(h->eos_dev.p, eos_ids_host, (size_t)n_eos * 8, cudaMemcpyHostToDevice, st));
    p.eos_ids = h->eos_dev.as<long long>();
    // ---- device state block
    int host_state[PADT_STATE_INTS] = {0};
    host_state[0] = 0; host_state[1] = B; host_state[2] = L;
    for (int b = 0; b < B; ++b) {
        PADT_REQUIRE(kv_start_host[b] >= 0 && kv_start_host[b] < L, "kv_start");
        PADT_REQUIRE(vrt_lo_host[b] >= 0 && vrt_hi_host[b] <= n_protos && vrt_lo_host[b] <= vrt_hi_host[b], "VRT range");
        host_state[PADT_STATE_HDR + 0 * PADT_STATE_BMAX + b] = kv_start_host[b];
        host_state[PADT_STATE_HDR + 1 * PADT_STATE_BMAX + b] = rope_delta_host[b];
        host_state[PADT_STATE_HDR + 2 * PADT_STATE_BMAX + b] = vrt_lo_host[b];
        host_state[PADT_STATE_HDR + 3 * PADT_STATE_BMAX + b] = vrt_hi_host[b];
        host_state[PADT_STATE_HDR + 4 * PADT_STATE_BMAX + b] = 1;
    }
    PADT_CHECK_CUDA(cudaMemcpyAsync(p.dev_state, host_state, sizeof(host_state), cudaMemcpyHostToDevice, st));
    // ---- segment tables for the causal prefill attention: one segment per row, keys in the cache of that row
    std::vector<int> seg(5 * B + 1, 0);
    for (int b = 0; b < B; ++b) {
        seg[0 * B + b] = b * L;                      // q_off (rows of the [T, qkv_n] projection buffer)
        seg[1 * B + b] = L;                          // q_len
        seg[2 * B + b] = b * nkv * h->s_max;         // k_off (rows of hd elements inside this layer's cache)
        seg[3 * B + b] = L;                          // k_len
        seg[4 * B + b] = kv_start_host[b];           // k_start (left pads are never attended)
    }
    RC(h->seg_ints.ensure(seg.size() * sizeof(int) + 4));
    int* segd = h->seg_ints.as<int>();
    PADT_CHECK_CUDA(cudaMemcpyAsync(segd, seg.data(), seg.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    int* err_flag = segd + 5 * B;
    PADT_CHECK_CUDA(cudaStreamSynchronize(st));   // host_state / seg are stack/pageable memory
    // ---- workspace: hcur [T,D] | n [T,D] | qkv [T,qkv_n] | attn [T,nh*hd] | act [T,I]
    size_t e_h = (size_t)T * D, e_qkv = (size_t)T * qkv_n, e_attn = (size_t)T * nh * hd, e_act = (size_t)T * I;
    RC(h->llm_ws.ensure((2 * e_h + e_qkv + e_attn + e_act) * sizeof(bf16)));
    bf16* x = h->llm_ws.as<bf16>();
    bf16* n = x + e_h;
    bf16* qkv = n + e_h;
    bf16* attn = qkv + e_qkv;
    bf16* act = attn + e_attn;
    // inputs_embeds = ext[input_ids] with image features scattered in (padt.py:204-219)
    RC(launch_embed_gather((const long long*)input_ids, T, h->embed, d.vocab, (const bf16*)prototypes, n_protos, (const bf16*)image_embeds, image_slot, D, x,
                           err_flag, st));
    const float scale = 1.0f / sqrtf((float)hd);
    for (int li = 0; li < d.n_layers; ++li) {
        const DecodeLayerW& w = h->llm[li];
        bf16* kc = p.kcache + (size_t)li * layer_kv;
        bf16* vc = p.vcache + (size_t)li * layer_kv;
        RC(launch_rmsnorm(x, D, w.ln1, n, D, T, D, d.rms_eps, st));
        RC(gemm(EPI_NONE, n, D, w.wqkv, D, w.bqkv, nullptr, 0, qkv, qkv_n, nullptr, T, qkv_n, D, st));
        RC(launch_mrope_kv(qkv, T, nh, nkv, hd, pos3, h->cos_tab, h->sin_tab, h->max_pos, d.mrope0, d.mrope1, kc, vc, B, L, h->s_max, st));
        AttnParams a = {};
        a.q = qkv; a.k = kc; a.v = vc; a.o = attn;
        a.q_row = qkv_n; a.q_head = hd; a.k_row = hd; a.k_head = (long long)h->s_max * hd; a.v_row = hd; a.v_head = (long long)h->s_max * hd;
        a.o_row = nh * hd; a.o_head = hd;
        a.q_off = segd; a.q_len = segd + B; a.k_off = segd + 2 * B; a.k_len = segd + 3 * B; a.k_start = segd + 4 * B;
        a.n_seg = B; a.n_heads = nh; a.group = nh / nkv; a.max_q_len = L; a.head_dim = hd; a.causal = 1; a.scale = scale;
        const AttnSegHost hseg = {seg.data(), seg.data() + B, seg.data() + 2 * B, seg.data() + 3 * B};
        a.host_seg = &hseg;
        RC(launch_attention(a, st));
        RC(gemm(EPI_RESID, attn, nh * hd, w.wo, nh * hd, nullptr, x, D, x, D, nullptr, T, D, nh * hd, st));
        RC(launch_rmsnorm(x, D, w.ln2, n, D, T, D, d.rms_eps, st));
        RC(gemm(EPI_SWIGLU, n, D, w.wgu, D, nullptr, nullptr, 0, act, I, nullptr, T, 2 * I, D, st));
        RC(gemm(EPI_RESID, act, I, w.wdown, I, nullptr, x, D, x, D, nullptr, T, D, I, st));
    }
    if (bufs->prefill_hidden) RC(launch_rmsnorm(x, D, h->final_norm, (bf16*)bufs->prefill_hidden, D, T, D, d.rms_eps, st));
    // last position of every row -> decode residual buffer, then final norm + extended logits + arg-max partials
    RC(launch_copy_rows_strided(x + (size_t)(L - 1) * D, (long long)L * D, p.h, D, B, D, st));
    RC(launch_final_logits(p, st));
    int err = 0;
    PADT_CHECK_CUDA(cudaMemcpyAsync(&err, err_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PADT_CHECK_CUDA(cudaStreamSynchronize(st));
    PADT_REQUIRE(err == 0, "input_ids contains an id outside [0, vocab + n_prototypes)  (padt.py:203)");
    h->gen_ready = true;
    return 0;
}

extern "C" int padt_decode(PadtHandle* h, int n_steps, int finish, padt_stream_t stream) {
    PADT_REQUIRE(h && h->gen_ready, "padt_prefill must run first");
    cudaStream_t caller = (cudaStream_t)stream;
    // Stream capture is illegal on the legacy default stream (which is what torch hands out by default), so the decode loop
    // runs on an engine-owned non-blocking stream, ordered after / before the caller's stream with events.
    if (!h->own_stream) {
        PADT_CHECK_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
        PADT_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
        PADT_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    cudaStream_t st = h->own_stream;
    PADT_CHECK_CUDA(cudaEventRecord(h->ev_in, caller));
    PADT_CHECK_CUDA(cudaStreamWaitEvent(st, h->ev_in, 0));
    int s = 0;
    if (n_steps > 0 && h->use_graph && !h->step_graph) {
        // first step runs eagerly (also performs the one-time cudaFuncSetAttribute calls), later steps replay a graph of it;
        // every step-dependent quantity lives in the device state block, so the replay is exact
        RC(run_step(h, st));
        s = 1;
        if (n_steps > 1) {
            cudaGraph_t g;
            PADT_CHECK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            unsigned long long before = g_padt_launches;
            int rc = run_step(h, st);
            h->graph_kernels = (int)(g_padt_launches - before);
            g_padt_launches = before;   // captured, not executed
            cudaError_t ce = cudaStreamEndCapture(st, &g);
            if (rc) return rc;
            PADT_CHECK_CUDA(ce);
            PADT_CHECK_CUDA(cudaGraphInstantiate(&h->step_graph, g, 0));
            cudaGraphDestroy(g);
        }
    }
    for (; s < n_steps; ++s) {
        if (h->use_graph) { PADT_CHECK_CUDA(cudaGraphLaunch(h->step_graph, st)); g_padt_launches += h->graph_kernels; }
        else RC(run_step(h, st));
    }
    if (finish) RC(launch_sample_embed(h->dp, st));
    PADT_CHECK_CUDA(cudaEventRecord(h->ev_out, st));
    PADT_CHECK_CUDA(cudaStreamWaitEvent(caller, h->ev_out, 0));
    return 0;
}

extern "C" int padt_decode_status(PadtHandle* h, int32_t* n_sampled, int32_t* n_unfinished, padt_stream_t stream) {
    PADT_REQUIRE(h && h->gen_ready, "padt_prefill must run first");
    cudaStream_t st = (cudaStream_t)stream;
    int v[2] = {0, 0};
    PADT_CHECK_CUDA(cudaMemcpyAsync(v, h->dp.dev_state, sizeof(v), cudaMemcpyDeviceToHost, st));
    PADT_CHECK_CUDA(cudaStreamSynchronize(st));
    if (n_sampled) *n_sampled = v[0];
    if (n_unfinished) *n_unfinished = v[1];
    return 0;
}
