// ViT patch encoder + VRT prototype projection (host orchestration).
// Reference: custom_visual_forward padt.py:48-106 (HF:91-146,207-322,382-453) and padt.py:187-191.
#include "engine.h"
#include <algorithm>

namespace {

struct HostPlan {
    std::vector<int> embed_row_map, hpos, wpos, merge_row_map, win_off, win_len, full_off, full_len;
};

// rot_pos_emb (HF:382-404) + get_window_index (HF:411-453) + cu_seqlens (padt.py:79-87), all integer host work.
void build_host_plan(const PadtDims& d, const int64_t* grid, int n_img, HostPlan& hp, int& P, int& M) {
    const int merge = d.vit_merge, unit = merge * merge;
    const int vws = d.vit_window / merge / d.vit_patch;
    std::vector<int> hpos_raw, wpos_raw, window_index;
    int unit_base = 0, patch_base = 0;
    for (int im = 0; im < n_img; ++im) {
        int t = (int)grid[im * 3], hh = (int)grid[im * 3 + 1], ww = (int)grid[im * 3 + 2];
        for (int ti = 0; ti < t; ++ti)
            for (int by = 0; by < hh / merge; ++by)
                for (int bx = 0; bx < ww / merge; ++bx)
                    for (int iy = 0; iy < merge; ++iy)
                        for (int ix = 0; ix < merge; ++ix) { hpos_raw.push_back(by * merge + iy); wpos_raw.push_back(bx * merge + ix); }
        int lh = hh / merge, lw = ww / merge;
        int pad_h = vws - lh % vws, pad_w = vws - lw % vws;   // a full extra (empty) window when divisible, as in HF
        int nwh = (lh + pad_h) / vws, nww = (lw + pad_w) / vws;
        for (int ti = 0; ti < t; ++ti)
            for (int wy = 0; wy < nwh; ++wy)
                for (int wx = 0; wx < nww; ++wx) {
                    int cnt = 0;
                    for (int iy = 0; iy < vws; ++iy)
                        for (int ix = 0; ix < vws; ++ix) {
                            int y = wy * vws + iy, x = wx * vws + ix;
                            if (y < lh && x < lw) { window_index.push_back(unit_base + ti * lh * lw + y * lw + x); ++cnt; }
                        }
                    if (cnt > 0) {   // torch.unique_consecutive drops the empty windows (padt.py:67)
                        int off = hp.win_off.empty() ? 0 : hp.win_off.back() + hp.win_len.back();
                        hp.win_off.push_back(off);
                        hp.win_len.push_back(cnt * unit);
                    }
                }
        for (int ti = 0; ti < t; ++ti) { hp.full_off.push_back(patch_base + ti * hh * ww); hp.full_len.push_back(hh * ww); }
        unit_base += t * lh * lw;
        patch_base += t * hh * ww;
    }
    P = patch_base;
    M = unit_base;
    std::vector<int> inv(M);
    for (int i = 0; i < M; ++i) inv[window_index[i]] = i;
    hp.embed_row_map.resize(P); hp.hpos.resize(P); hp.wpos.resize(P);
    for (int r = 0; r < P; ++r) hp.embed_row_map[r] = inv[r / unit] * unit + r % unit;     // padt.py:70-72 as a scatter
    for (int i = 0; i < M; ++i)
        for (int k = 0; k < unit; ++k) {
            hp.hpos[i * unit + k] = hpos_raw[window_index[i] * unit + k];                   // padt.py:73-75
            hp.wpos[i * unit + k] = wpos_raw[window_index[i] * unit + k];
        }
    hp.merge_row_map = window_index;                                                        // padt.py:103-104 as a scatter
}

int get_plan(PadtHandle* h, const int64_t* grid, int n_img, VitPlan** out, cudaStream_t st) {
    std::string key;
    for (int i = 0; i < n_img * 3; ++i) key += std::to_string(grid[i]) + ",";
    auto it = h->vit_plans.find(key);
    if (it != h->vit_plans.end()) {
        // most recently used key moves to the back of the LRU order
        for (size_t i = 0; i < h->vit_plan_order.size(); ++i)
            if (h->vit_plan_order[i] == key) { h->vit_plan_order.erase(h->vit_plan_order.begin() + i); break; }
        h->vit_plan_order.push_back(key);
        *out = it->second;
        return 0;
    }
    HostPlan hp;
    int P, M;
    build_host_plan(h->d, grid, n_img, hp, P, M);
    VitPlan* pl = new VitPlan();
    pl->P = P; pl->M = M; pl->n_win = (int)hp.win_off.size(); pl->n_full = (int)hp.full_off.size();
    for (int v : hp.win_len) pl->max_win = v > pl->max_win ? v : pl->max_win;
    for (int v : hp.full_len) pl->max_full = v > pl->max_full ? v : pl->max_full;
    std::vector<int> packed;
    auto put = [&](const std::vector<int>& v) { size_t off = packed.size(); packed.insert(packed.end(), v.begin(), v.end()); while (packed.size() % 4) packed.push_back(0); return off; };
    size_t o0 = put(hp.embed_row_map), o1 = put(hp.hpos), o2 = put(hp.wpos), o3 = put(hp.merge_row_map);
    size_t o4 = put(hp.win_off), o5 = put(hp.win_len), o6 = put(hp.full_off), o7 = put(hp.full_len);
    int rc = pl->ints.ensure(packed.size() * sizeof(int));
    if (rc) { delete pl; return rc; }
    cudaError_t ce = cudaMemcpyAsync(pl->ints.p, packed.data(), packed.size() * sizeof(int), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);   // `packed` is pageable host memory
    if (ce != cudaSuccess) {                                   // do not leak the half-built plan
        pl->ints.release();
        delete pl;
        padt_set_error(std::string("ViT plan upload: ") + cudaGetErrorString(ce));
        return -2;
    }
    // Bounded cache (an eval run over images of many sizes makes nearly every batch a new key): evict the least recently used plan.
    // Safe here: the stream was just synchronised, so no launch in flight still reads an old plan's device tables.
    while ((int)h->vit_plan_order.size() >= PADT_VIT_PLAN_CACHE) {
        auto old = h->vit_plans.find(h->vit_plan_order.front());
        if (old != h->vit_plans.end()) { old->second->ints.release(); delete old->second; h->vit_plans.erase(old); }
        h->vit_plan_order.erase(h->vit_plan_order.begin());
    }
    h->vit_plan_order.push_back(key);
    int* b = pl->ints.as<int>();
    pl->embed_row_map = b + o0; pl->hpos = b + o1; pl->wpos = b + o2; pl->merge_row_map = b + o3;
    pl->win_off = b + o4; pl->win_len = b + o5; pl->full_off = b + o6; pl->full_len = b + o7;
    pl->full_off_host = hp.full_off; pl->full_len_host = hp.full_len;
    h->vit_plans[key] = pl;
    *out = pl;
    return 0;
}

}  // namespace

// Host-only: the integer tables of V3/V4 (window permutation, rotary position ids, segment lengths) for one grid list.
// Sizes: embed_row_map/hpos/wpos [P], merge_row_map [M], win_len [<= M]; returns the number of windows via *n_win.
extern "C" int padt_vit_plan_host(const PadtDims* dims, const int64_t* grid_thw_host, int n_img, int32_t* embed_row_map, int32_t* hpos,
                                  int32_t* wpos, int32_t* merge_row_map, int32_t* win_len, int32_t* n_win, int32_t* n_patches) {
    PADT_REQUIRE(dims && grid_thw_host && n_win && n_patches, "null argument");
    HostPlan hp;
    int P, M;
    build_host_plan(*dims, grid_thw_host, n_img, hp, P, M);
    *n_patches = P;
    *n_win = (int)hp.win_len.size();
    if (embed_row_map) std::copy(hp.embed_row_map.begin(), hp.embed_row_map.end(), embed_row_map);
    if (hpos) std::copy(hp.hpos.begin(), hp.hpos.end(), hpos);
    if (wpos) std::copy(hp.wpos.begin(), hp.wpos.end(), wpos);
    if (merge_row_map) std::copy(hp.merge_row_map.begin(), hp.merge_row_map.end(), merge_row_map);
    if (win_len) std::copy(hp.win_len.begin(), hp.win_len.end(), win_len);
    return 0;
}

extern "C" int padt_vit_forward(PadtHandle* h, const void* pixel_values, const int64_t* grid_thw_host, int n_img, const float* pos_cos,
                                const float* pos_sin, int tab_rows, void* image_embeds, void* high_res, float* cos, float* sin,
                                padt_stream_t stream) {
    PADT_REQUIRE(h && pixel_values && grid_thw_host && image_embeds && high_res && cos && sin, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const PadtDims& d = h->d;
    int max_grid = 0;
    for (int i = 0; i < n_img; ++i) {
        int hh = (int)grid_thw_host[i * 3 + 1], ww = (int)grid_thw_host[i * 3 + 2];
        PADT_REQUIRE(hh % d.vit_merge == 0 && ww % d.vit_merge == 0 && grid_thw_host[i * 3] >= 1, "grid must be a multiple of the merge size");
        max_grid = hh > max_grid ? hh : max_grid;
        max_grid = ww > max_grid ? ww : max_grid;
    }
    PADT_REQUIRE(max_grid <= tab_rows, "rotary position table too short for this grid");
    VitPlan* pl;
    RC(get_plan(h, grid_thw_host, n_img, &pl, st));
    const int P = pl->P, M = pl->M, Hv = d.vit_hidden, Ip = h->vit_ipad, hd = h->vit_hd, nh = d.vit_heads;
    const int Kpe = d.vit_in_ch * d.vit_tpatch * d.vit_patch * d.vit_patch;
    PADT_REQUIRE(Kpe % 8 == 0, "patch dim");
    const int unit = d.vit_merge * d.vit_merge;
    // workspace: n [P,Hv], qkv [P,3Hv], attn [P,Hv], act [P,Ip], mg [M, unit*Hv]
    size_t e_n = (size_t)P * Hv, e_qkv = (size_t)P * 3 * Hv, e_act = (size_t)P * Ip, e_mg = (size_t)M * unit * Hv;
    RC(h->vit_ws.ensure((e_n * 2 + e_qkv + e_act + e_mg) * sizeof(bf16)));
    bf16* n = h->vit_ws.as<bf16>();
    bf16* qkv = n + e_n;
    bf16* attn = qkv + e_qkv;
    bf16* act = attn + e_n;
    bf16* mg = act + e_act;
    bf16* x = reinterpret_cast<bf16*>(high_res);   // the residual stream lives in the caller's high_res buffer

    // patch embed (HF:108-114: Conv3d with stride == kernel is a GEMM over [P,1176]); rows stored in window order (padt.py:70-72)
    RC(gemm(EPI_NONE, (const bf16*)pixel_values, Kpe, h->w_patch, Kpe, nullptr, nullptr, 0, x, Hv, pl->embed_row_map, P, Hv, Kpe, st));
    RC(launch_vit_cos_sin(pos_cos, pos_sin, hd / 4, pl->hpos, pl->wpos, cos, sin, P, st));
    const float scale = 1.0f / sqrtf((float)hd);
    for (int li = 0; li < d.vit_depth; ++li) {
        const VitBlockW& b = h->vit[li];
        bool full = false;
        for (int k = 0; k < d.vit_n_fullatt; ++k) full |= (d.vit_fullatt[k] == li);
        RC(launch_rmsnorm(x, Hv, b.norm1, n, Hv, P, Hv, 1e-6f, st));
        RC(gemm(EPI_NONE, n, Hv, b.wqkv, Hv, b.bqkv, nullptr, 0, qkv, 3 * Hv, nullptr, P, 3 * Hv, Hv, st));
        RC(launch_rope_neox_f32(qkv, 3 * Hv, 2 * nh, hd, cos, sin, hd, nullptr, P, st));    // q heads then k heads are contiguous
        AttnParams a = {};
        a.q = qkv; a.k = qkv + Hv; a.v = qkv + 2 * Hv; a.o = attn;
        a.q_row = a.k_row = a.v_row = 3 * Hv; a.q_head = a.k_head = a.v_head = hd; a.o_row = Hv; a.o_head = hd;
        a.q_off = a.k_off = full ? pl->full_off : pl->win_off;
        a.q_len = a.k_len = full ? pl->full_len : pl->win_len;
        a.k_start = nullptr;
        a.n_seg = full ? pl->n_full : pl->n_win; a.n_heads = nh; a.group = 1; a.max_q_len = full ? pl->max_full : pl->max_win;
        a.head_dim = hd; a.causal = 0; a.scale = scale;
        const AttnSegHost hseg = {pl->full_off_host.data(), pl->full_len_host.data(), pl->full_off_host.data(), pl->full_len_host.data()};
        if (full) a.host_seg = &hseg;
        RC(launch_attention(a, st));
        RC(gemm(EPI_RESID, attn, Hv, b.wproj, Hv, b.bproj, x, Hv, x, Hv, nullptr, P, Hv, Hv, st));
        RC(launch_rmsnorm(x, Hv, b.norm2, n, Hv, P, Hv, 1e-6f, st));
        RC(gemm(EPI_SWIGLU, n, Hv, b.wgu, Hv, b.bgu, nullptr, 0, act, Ip, nullptr, P, 2 * d.vit_inter, Hv, st));
        RC(gemm(EPI_RESID, act, Ip, b.wdown, Ip, b.bdown, x, Hv, x, Hv, nullptr, P, Hv, d.vit_inter, st));
    }
    // merger (HF:144-146): RMSNorm -> [M, 4*Hv] view -> Linear+GELU -> Linear; rows un-permuted to raster order (padt.py:103-104)
    RC(launch_rmsnorm(x, Hv, h->mg_ln, n, Hv, P, Hv, 1e-6f, st));
    RC(gemm(EPI_GELU, n, unit * Hv, h->mg_w0, unit * Hv, h->mg_b0, nullptr, 0, mg, unit * Hv, nullptr, M, unit * Hv, unit * Hv, st));
    RC(gemm(EPI_NONE, mg, unit * Hv, h->mg_w2, unit * Hv, h->mg_b2, nullptr, 0, (bf16*)image_embeds, d.vit_out_hidden, pl->merge_row_map, M,
            d.vit_out_hidden, unit * Hv, st));
    return 0;
}

extern "C" int padt_prototypes(PadtHandle* h, const void* image_embeds, int n_rows, void* prototypes, padt_stream_t stream) {
    PADT_REQUIRE(h && image_embeds && prototypes, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const PadtDims& d = h->d;
    const int D = d.D;
    if (!d.use_proto_proj) {   // padt.py:191
        PADT_CHECK_CUDA(cudaMemcpyAsync(prototypes, image_embeds, (size_t)n_rows * D * 2, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    RC(h->proto_ws.ensure((size_t)n_rows * d.lora_r * 2));
    bf16* t = h->proto_ws.as<bf16>();
    bf16* p = reinterpret_cast<bf16*>(prototypes);
    RC(launch_layernorm((const bf16*)image_embeds, h->vn_w, h->vn_b, p, n_rows, D, 1e-5f, st));     // padt.py:188
    RC(gemm(EPI_NONE, p, D, h->vp_w0, D, nullptr, nullptr, 0, t, d.lora_r, nullptr, n_rows, d.lora_r, D, st));
    RC(gemm(EPI_RESID, t, d.lora_r, h->vp_w1, d.lora_r, nullptr, p, D, p, D, nullptr, n_rows, D, d.lora_r, st));  // padt.py:189
    return 0;
}
