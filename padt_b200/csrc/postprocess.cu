// Result post-processing of the batched inference harness on the GPU (eval/evaluation_scripts/utils.py:260-264):
//   mask = F.interpolate(pred_mask[:4H, :4W], size = (h, w), mode = 'bilinear').sigmoid() > 0.5  ->  pycocotools RLE.
// One kernel evaluates the bilinear sample (ATen upsample_bilinear2d arithmetic: fp32, align_corners = False), rounds it to bf16 as the
// reference's bf16 tensor does, and thresholds: a bf16 x has bf16(sigmoid(x)) > 0.5 exactly when x > 2^-7 (sigmoid(2^-7) = 0.501953115
// rounds down to 0.5, the next bf16 rounds up to 0.50390625).  The mask is written column-major (what np.asfortranarray hands to
// pycocotools) and the run starts of its column-major flattening are compacted on the device (cub::DeviceSelect), so only the runs, not
// h * w bytes per object, cross PCIe.  HBM-bound byte work; no tensor cores.
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include "common.cuh"
#include "kernels.h"
#include "../../include/padt_b200.h"

namespace {

__global__ void mask_threshold_kernel(const bf16* __restrict__ logits, long long row_stride, int in_h, int in_w, int out_h, int out_w,
                                      uint8_t* __restrict__ dense_cm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // column-major: i = x * out_h + y
    if (i >= (long long)out_h * out_w) return;
    const int w2 = (int)(i / out_h), h2 = (int)(i % out_h);
    const float rheight = (float)in_h / (float)out_h, rwidth = (float)in_w / (float)out_w;
    // area_pixel_compute_source_index(scale, dst, align_corners = false, cubic = false)
    float h1r = rheight * ((float)h2 + 0.5f) - 0.5f;
    h1r = h1r < 0.f ? 0.f : h1r;
    const int h1 = (int)h1r;
    const int h1p = (h1 < in_h - 1) ? 1 : 0;
    const float h1lambda = h1r - (float)h1, h0lambda = 1.0f - h1lambda;
    float w1r = rwidth * ((float)w2 + 0.5f) - 0.5f;
    w1r = w1r < 0.f ? 0.f : w1r;
    const int w1 = (int)w1r;
    const int w1p = (w1 < in_w - 1) ? 1 : 0;
    const float w1lambda = w1r - (float)w1, w0lambda = 1.0f - w1lambda;
    const bf16* r0 = logits + (long long)h1 * row_stride;
    const bf16* r1 = logits + (long long)(h1 + h1p) * row_stride;
    const float val = h0lambda * (w0lambda * __bfloat162float(r0[w1]) + w1lambda * __bfloat162float(r0[w1 + w1p])) +
                      h1lambda * (w0lambda * __bfloat162float(r1[w1]) + w1lambda * __bfloat162float(r1[w1 + w1p]));
    dense_cm[i] = __bfloat162float(__float2bfloat16_rn(val)) > 0.0078125f ? 1 : 0;
}

struct RunStart {
    const uint8_t* m;
    __device__ __forceinline__ bool operator()(const int& i) const { return m[i] != (i ? m[i - 1] : (uint8_t)0); }
};

size_t align256p(size_t x) { return (x + 255) & ~size_t(255); }

size_t select_temp_bytes(int n) {
    size_t bytes = 0;
    cub::CountingInputIterator<int> it(0);
    RunStart op = {nullptr};
    cub::DeviceSelect::If(nullptr, bytes, it, (int*)nullptr, (int*)nullptr, n, op, (cudaStream_t)0);
    return bytes;
}

}  // namespace

extern "C" size_t padt_mask_runs_workspace_bytes(int out_h, int out_w) {
    if (out_h <= 0 || out_w <= 0) return 0;
    return align256p(select_temp_bytes(out_h * out_w)) + 256;
}

extern "C" int padt_mask_to_runs(const void* mask_logits, int64_t row_stride, int valid_h, int valid_w, int out_h, int out_w, uint8_t* dense_cm,
                                 int32_t* run_starts, int32_t* n_runs, void* workspace, size_t workspace_bytes, padt_stream_t stream) {
    PADT_REQUIRE(mask_logits && dense_cm && run_starts && n_runs && workspace, "null pointer");
    PADT_REQUIRE(valid_h > 0 && valid_w > 0 && out_h > 0 && out_w > 0 && row_stride >= valid_w, "sizes");
    PADT_REQUIRE((long long)out_h * out_w < (1ll << 31), "image too large");
    PADT_REQUIRE(workspace_bytes >= padt_mask_runs_workspace_bytes(out_h, out_w), "workspace too small (padt_mask_runs_workspace_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = out_h * out_w;
    mask_threshold_kernel<<<cdiv(n, 256), 256, 0, st>>>((const bf16*)mask_logits, row_stride, valid_h, valid_w, out_h, out_w, dense_cm);
    PADT_LAUNCH_CHECK();
    size_t temp = workspace_bytes;
    cub::CountingInputIterator<int> it(0);
    RunStart op = {dense_cm};
    PADT_CHECK_CUDA(cub::DeviceSelect::If(workspace, temp, it, run_starts, n_runs, n, op, st));
    g_padt_launches += 2;
    return 0;
}

// pycocotools rleToString (cocoapi common/maskApi.c) from the run starts of the column-major mask: counts = differences of the starts
// (first count = zeros before the first 1, possibly 0), every count from the fourth on stored as the difference to the count two back,
// 5 bits per character, 0x20 = continuation, + 48.  Host function.  Returns the string length, or -1 when `cap` is too small.
extern "C" int padt_rle_string(const int32_t* run_starts_host, int n_starts, int total, char* out, int cap) {
    if (!out || cap <= 0 || n_starts < 0 || (n_starts > 0 && !run_starts_host)) { padt_set_error("padt_rle_string: arguments"); return -1; }
    int p = 0;
    long c1 = 0, c2 = 0;   // counts one and two back
    const int m = n_starts + 1;
    for (int i = 0; i < m; ++i) {
        const long lo = i == 0 ? 0 : run_starts_host[i - 1], hi = i < n_starts ? run_starts_host[i] : total;
        const long cnt = hi - lo;
        long x = cnt;
        if (i > 2) x -= c2;
        c2 = c1; c1 = cnt;
        bool more = true;
        while (more) {
            char c = (char)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more) c |= 0x20;
            c += 48;
            if (p + 1 >= cap) { padt_set_error("padt_rle_string: output buffer too small"); return -1; }
            out[p++] = c;
        }
    }
    out[p] = 0;
    return p;
}
