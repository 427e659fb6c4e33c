// Host-side pre/post-processing of the reference entry points, batched and on the device.
//
// Replaces, per image, infer_dataset_lol.py:124-128 / infer_unpaired.py:81-88,121-122 (reflect padding, /255, log(clamp(x + 1e-3, 1e-3)))
// and infer_unpaired.py:40-42,130 / infer_dataset_lol.py:135 (crop the padding, clip to [0,1], * 255, truncate to uint8).
//   preprocess : uint8 NHWC [B,H,W,3] -> fp32 NCHW [B,3,Hp,Wp]; source pixel of padded (y, x) by reflection:
//                mode 0 "reflect"  (np.pad 'reflect', no edge repeat:  -1 -> 1)    -- infer_dataset_lol.py impad
//                mode 1 "symmetric" (cv2.BORDER_REFLECT, edge repeated: -1 -> 0)   -- infer_unpaired.py auto_padding
//   postprocess: fp32 [B,3,Hp,Wp] with arbitrary element strides -> uint8 NHWC [B,H,W,3] of the crop box.
// HBM-bound streaming kernels (3 channels); one thread per output pixel.
#include "common.cuh"

namespace glare {

__device__ __forceinline__ int reflect_index(int i, int n, int mode) {
    if (mode == 0) {                      // reflect without repeating the edge
        if (i < 0) i = -i;
        if (i >= n) i = 2 * n - 2 - i;
    } else {                              // symmetric: edge pixel repeated
        if (i < 0) i = -i - 1;
        if (i >= n) i = 2 * n - 1 - i;
    }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ img, int B, int H, int W, int top, int left, int Hp,
                                                         int Wp, int mode, float* __restrict__ out) {
    const long long n = (long long)B * Hp * Wp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wp);
        long long r = i / Wp;
        const int y = (int)(r % Hp), b = (int)(r / Hp);
        const int sy = reflect_index(y - top, H, mode), sx = reflect_index(x - left, W, mode);
        const uint8_t* p = img + (((long long)b * H + sy) * W + sx) * 3;
        float* o = out + ((long long)b * 3 * Hp + y) * Wp + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = (float)p[c] / 255.0f;                            // t(): astype(float32) / 255
            o[(long long)c * Hp * Wp] = logf(fmaxf(v + 1e-3f, 1e-3f));       // log(clamp(x + 1e-3, min=1e-3))
        }
    }
}

__global__ void __launch_bounds__(256) postprocess_kernel(const float* __restrict__ y, long long sb, long long sc, long long sh, long long sw,
                                                          int B, int y0, int x0, int H, int W, uint8_t* __restrict__ out) {
    const long long n = (long long)B * H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        long long r = i / W;
        const int yy = (int)(r % H), b = (int)(r / H);
        const float* p = y + b * sb + (long long)(y0 + yy) * sh + (long long)(x0 + x) * sw;
        uint8_t* o = out + i * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = p[c * sc];
            v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;                          // np.clip(t, 0, 1) * 255
            o[c] = (uint8_t)v;                                               // .astype(np.uint8): truncation (NaN -> 0)
        }
    }
}

}  // namespace glare

using namespace glare;

GLARE_API int glare_preprocess_u8(const uint8_t* img_nhwc, int B, int H, int W, int pad_top, int pad_bottom, int pad_left, int pad_right,
                                  int mode, float* out_nchw, cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || pad_top < 0 || pad_bottom < 0 || pad_left < 0 || pad_right < 0 || (mode != 0 && mode != 1)) return GLARE_ERR_BAD_ARG;
    if (pad_top >= H + mode || pad_bottom >= H + mode || pad_left >= W + mode || pad_right >= W + mode) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!img_nhwc || !out_nchw) return GLARE_ERR_BAD_ARG;
    const int Hp = H + pad_top + pad_bottom, Wp = W + pad_left + pad_right;
    const long long n = (long long)B * Hp * Wp;
    const int grid = (int)((n + 255) / 256 < 148LL * 16 ? (n + 255) / 256 : 148LL * 16);
    preprocess_kernel<<<grid, 256, 0, stream>>>(img_nhwc, B, H, W, pad_top, pad_left, Hp, Wp, mode, out_nchw);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}

// y: fp32, logical [B,3,Hp,Wp] addressed with element strides (sb, sc, sh, sw); writes uint8 NHWC [B,H,W,3] of the box at (y0, x0)
GLARE_API int glare_postprocess_u8(const float* y, long long sb, long long sc, long long sh, long long sw, int B, int y0, int x0, int H, int W,
                                   uint8_t* out_nhwc, cudaStream_t stream) {
    if (B < 0 || H <= 0 || W <= 0 || y0 < 0 || x0 < 0) return GLARE_ERR_BAD_ARG;
    if (B == 0) return GLARE_OK;
    if (!y || !out_nhwc) return GLARE_ERR_BAD_ARG;
    const long long n = (long long)B * H * W;
    const int grid = (int)((n + 255) / 256 < 148LL * 16 ? (n + 255) / 256 : 148LL * 16);
    postprocess_kernel<<<grid, 256, 0, stream>>>(y, sb, sc, sh, sw, B, y0, x0, H, W, out_nhwc);
    GLARE_CHECK_LAUNCH();
    return GLARE_OK;
}
